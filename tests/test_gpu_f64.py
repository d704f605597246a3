"""float64 path matrices (last paragraph; VERDICT round 1, missing #6): seven float64 arrays select the
all-float64 prologue (NumPy's dtype flow for float64 inputs).  The reference-generated goldens f64_* (tests/golden/cases.py) are
replayed by test_gpu_golden.py; here: the selection rules, every kernel family on float64 inputs against the oracle, and the
by-products' dtypes."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset, oracle_kwargs_from_params

pytestmark = pytest.mark.gpu
KEYS = ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el")


def _f64(data, seed=1):
    rng = np.random.default_rng(seed)
    return {k: (v.astype(np.float64) * (1.0 + 1e-6 * rng.standard_normal(v.shape)) if k in KEYS else v) for k, v in data.items()}


def test_mixed_dtypes_are_refused():
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import scenario
    s = scenario(1, 50)
    d = dict(s.data)
    d["delay"] = d["delay"].astype(np.float64)
    with pytest.raises(TypeError):
        dmb.Dataset(d).compute_channels(dmb.ChannelGenParameters(s.params), warn=False)


@pytest.mark.parametrize("cfg,n,variant", [(1, 400, "small"), (1, 400, "small1"), (1, 400, "mma"), (1, 400, "auto"), (6, 60, "auto"), (2, 10, "tc"), (2, 10, "tc1"), (2, 10, "ffma"), (3, 64, "auto"),
                                           (5, 40, "tile"), (4, 200, "auto")])
def test_float64_inputs_match_oracle_in_every_kernel_family(cfg, n, variant, monkeypatch):
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import scenario
    from oracle import channel_oracle as orc
    if variant != "auto":
        monkeypatch.setenv("DMK_FD_KERNEL", variant)
    s = scenario(cfg, n)
    d = _f64(s.data, cfg)
    H, info = make_dataset(dmb, d, s.bs_fov, s.ue_fov).compute_channels(dmb.ChannelGenParameters(s.params), times=s.times,
                                                                       doppler=s.doppler_hz, return_info=True, warn=False)
    o = orc.compute_channels(d, **oracle_kwargs_from_params(s.params, s.bs_fov, s.ue_fov), doppler_hz=s.doppler_hz, times=s.times)
    err = assert_channels_close(H, o["H"], what=f"{s.name} float64 {variant}")
    assert np.array_equal(info.valid, o["valid"])
    if o["fov_mask"] is not None:
        assert np.array_equal(info.fov_mask, o["fov_mask"])
    if s.params["freq_domain"]:
        assert np.array_equal(info.clip, o["clip"])
    print(f"{s.name} float64 [{info.kernel.split(' ')[0]}]: max per-user rel. Frobenius {err:.2e}")


def test_float64_byproducts_and_sionna_tau():
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import scenario
    from oracle import channel_oracle as orc
    s = scenario(3, 120)
    d = _f64(s.data, 9)
    ds = make_dataset(dmb, d, s.bs_fov, s.ue_fov)
    ds.set_channel_params(dmb.ChannelGenParameters(s.params))
    th, ph = orc.rotate_angles(np.asarray(s.params["bs_antenna"]["rotation"]), d["aod_el"], d["aod_az"])
    ok = ~np.isnan(th)
    np.testing.assert_allclose(ds["_aod_el_rot"][ok], th[ok], rtol=0, atol=1e-12)
    o = orc.compute_channels(d, **oracle_kwargs_from_params(s.params, s.bs_fov, s.ue_fov))
    assert np.array_equal(ds["_fov_mask"], o["fov_mask"])
    assert ds["_power_linear_ant_gain"].dtype == np.float64
    assert np.array_equal(ds.num_paths, o["fov_mask"].sum(1))
    p_lin = 10 ** (d["power"] / 10)
    tot = np.abs(np.nansum(np.sqrt(p_lin).astype(np.complex64) * np.exp(1j * np.deg2rad(d["phase"])), axis=1)) ** 2
    pl = ds.pl
    np.testing.assert_allclose(pl[tot > 0], -10 * np.log10(tot[tot > 0]), rtol=2e-6, atol=2e-5)
    from deepmimo_b200.sionna_adapter import DeepMIMOSionnaAdapter
    s4 = scenario(4, 60)
    d4 = _f64(s4.data, 4)
    a, tau = DeepMIMOSionnaAdapter(dmb.Dataset(d4), dmb.ChannelGenParameters(s4.params)).arrays()
    valid = ~np.isnan(d4["power"])
    for i in range(60):
        want = d4["delay"][i][valid[i]].astype(np.float32)
        assert np.array_equal(tau[i, 0, 0, :len(want)], want) and np.all(tau[i, 0, 0, len(want):] == 0)
