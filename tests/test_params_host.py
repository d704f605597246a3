"""Host logic: ChannelGenParameters / DotDict semantics (channel.py:20-139, general_utils.py:124-261),
parse_spec, Dataset mirror cache rules, sharding plan.  No GPU."""
import numpy as np
import pytest

import deepmimo_b200 as dmb
from deepmimo_b200.channels import delay_overflow_warning, parse_spec, resolve_ue_rotation
from deepmimo_b200.sharding import ShardItem, shard_plan


def test_defaults_match_reference_defaults():
    p = dmb.ChannelGenParameters()
    assert list(p.bs_antenna.shape) == [8, 1] and list(p.ue_antenna.shape) == [1, 1]
    assert p.bs_antenna.spacing == 0.5 and p["ue_antenna"]["radiation_pattern"] == "isotropic"
    assert p.num_paths == 25 and p.freq_domain == 1 and p.enable_doppler == 0 and p.enable_dual_polar == 0
    assert p.ofdm.subcarriers == 512 and p.ofdm.bandwidth == 10e6 and p.ofdm.rx_filter == 0
    assert np.array_equal(p.ofdm.selected_subcarriers, np.arange(1))
    assert dmb.ChannelParameters is dmb.ChannelGenParameters


def test_dotdict_semantics():
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array([4, 4])
    p["ofdm"]["subcarriers"] = 64
    assert p["bs_antenna"]["shape"][0] == 4 and p.ofdm.subcarriers == 64
    q = p.deepcopy()
    q.bs_antenna.shape[0] = 99
    assert p.bs_antenna.shape[0] == 4 and isinstance(q, dmb.ChannelGenParameters)
    p.update({"ofdm": {"subcarriers": 8}})                   # shallow, like the reference: sub-dict replaced
    assert "bandwidth" not in p.ofdm.keys()
    assert isinstance(p.to_dict()["ofdm"], dict)
    with pytest.raises(AttributeError):
        p.nope


def test_validate_assertions(capsys):
    p = dmb.ChannelGenParameters()
    p.bs_antenna.rotation = np.array([1, 2])
    with pytest.raises(AssertionError):
        p.validate(4)
    p = dmb.ChannelGenParameters()
    p.ue_antenna.rotation = np.zeros((5, 3))
    with pytest.raises(AssertionError):
        p.validate(4)
    p.validate(5)
    p = dmb.ChannelGenParameters()
    p.ue_antenna.radiation_pattern = "patch"
    with pytest.raises(AssertionError):
        p.validate(4)
    p = dmb.ChannelGenParameters()
    p.typo_key = 1
    p.validate(4)
    assert "typo_key" in capsys.readouterr().out
    p = dmb.ChannelGenParameters({"ue_antenna": {"shape": np.array([2, 1]), "spacing": 0.5}})
    p.validate(3)
    assert list(p.ue_antenna.rotation) == [0, 0, 0] and p.ue_antenna.radiation_pattern == "isotropic"


def test_parse_spec_fov_rules_and_subcarriers():
    p = dmb.ChannelGenParameters()
    p.ofdm.selected_subcarriers = np.arange(3) * 3
    s = parse_spec(p.validate(10), 10)
    assert (s.subc_start, s.subc_step) == (0, 3) and not s.fov_any and s.out_shape(10, 25) == (10, 1, 8, 3)
    p.ofdm.selected_subcarriers = np.array([0, 1, 5])
    assert parse_spec(p, 10).subc_step == 0
    s = parse_spec(p, 10, bs_fov=np.array([360, 180]), ue_fov=np.array([360, 180]))
    assert not s.fov_any                                      # dataset.py:484: both full -> mask is None
    s = parse_spec(p, 10, bs_fov=np.array([140, 120]), ue_fov=np.array([360, 180]))
    assert s.fov_any and s.fov_side == (True, False)
    p.ue_antenna.fov = np.array([90, 80])                     # alternative spelling inside the antenna dict
    s = parse_spec(p, 10)
    assert s.fov_any and s.fov_side == (False, True)
    p.freq_domain = 0
    p.num_paths = 10
    assert parse_spec(p, 10, times=[0, 1e-3]).out_shape(10, 25) == (10, 1, 8, 10, 2)
    p.ofdm.rx_filter = 1                                      # receive LPF (channel.py:193-194) is carried to the kernel
    p.enable_dual_polar = 1                                   # present in the defaults, read by nothing in the reference
    assert parse_spec(p, 10).rx_filter == 1


def test_random_ue_rotation_matches_reference_draw():
    """dataset.py:250,:332-338: seed 1001 then U(lo, hi) per user from NumPy's global RNG."""
    rng = np.array([[0, 30], [-20, 20], [0, 90]])
    _, got = resolve_ue_rotation(rng, 7)
    np.random.seed(1001)
    want = np.random.uniform(rng[:, 0], rng[:, 1], (7, 3))
    assert np.array_equal(got, want)
    from oracle import channel_oracle as orc
    assert np.array_equal(orc.resolve_ue_rotation(rng, 7), want)
    u, per = resolve_ue_rotation(np.array([1, 2, 3]), 7)
    assert per is None and list(u) == [1, 2, 3]
    with pytest.raises(ValueError):
        resolve_ue_rotation(np.zeros((6, 3)), 7)


def test_delay_overflow_warning(capsys):
    p = dmb.ChannelGenParameters()
    s = parse_spec(p.validate(2), 2)
    d = np.array([[1e-6, np.nan], [60e-6, np.nan]], np.float32)
    assert delay_overflow_warning(d, s, 2)
    assert "exceed OFDM symbol duration" in capsys.readouterr().out
    assert not delay_overflow_warning(d[:1], s, 2)
    assert not delay_overflow_warning(np.full((2, 2), np.nan, np.float32), s, 2)


def test_dataset_mirror_cache_rules(monkeypatch):
    calls = []

    def fake(ds, params, **kw):
        calls.append(params)
        H = np.zeros((ds.n_ue, 1, 8, 1), np.complex64)
        ds["channel"] = H
        return H

    import deepmimo_b200.dataset as dsmod
    monkeypatch.setattr(dsmod._ch, "compute_channels", fake)
    data = {"power": np.zeros((3, 25), np.float32), "rx_pos": np.zeros((3, 3), np.float32)}
    ds = dmb.Dataset(data)
    assert ds.n_ue == 3 and ds["pwr"] is data["power"]
    H = ds.channel                                            # lazy key triggers compute_channels (dataset.py:837)
    assert H.shape == (3, 1, 8, 1) and len(calls) == 1 and ds.ch is H
    ds.apply_fov(bs_fov=np.array([120, 90]))
    assert "channel" not in ds.keys() and list(ds.bs_fov) == [120, 90] and list(ds.ue_fov) == [360, 180]
    p = dmb.ChannelGenParameters()
    ds.set_channel_params(p)
    ds._data["_aod_el_rot"] = 1
    p2 = dmb.ChannelGenParameters(); p2.bs_antenna.rotation = np.array([0, 0, 5])
    ds.set_channel_params(p2)                                 # rotation changed -> rotated-angle cache cleared (dataset.py:214-220)
    assert "_aod_el_rot" not in ds.keys()
    macro = dmb.MacroDataset([dmb.Dataset(dict(data)), dmb.Dataset(dict(data))])
    res = macro.compute_channels(p)
    assert isinstance(res, list) and len(res) == 2 and len(macro) == 2
    assert dmb.MacroDataset([ds]).compute_channels(p).shape == (3, 1, 8, 1)


def test_shard_plan_covers_every_user_once():
    for sizes, w in (([10], 3), ([5, 7, 3], 4), ([200_000] * 8, 8), ([3], 8), ([0, 9], 2)):
        plan = shard_plan(sizes, w)
        assert len(plan) == w
        seen = [np.zeros(n, int) for n in sizes]
        for items in plan:
            for it in items:
                seen[it.bs][it.start:it.stop] += 1
        assert all((s == 1).all() for s in seen)
        loads = [sum(i.n for i in items) for items in plan]
        assert max(loads) - min(loads) <= 1
    assert shard_plan([200_000] * 8, 8)[3] == [ShardItem(3, 0, 200_000)]
