"""Row a11 pinned: the v3 generator inside the reference tree is the only Doppler definition (deepmimo_v3/generator/python/
construct_deepmimo.py:267-280).  tests/golden/doppler_v3.npz holds the channel that code produced on seeded ray data
(tests/golden/make_golden_doppler_v3.py); the oracle and the CUDA path must reproduce it through `enable_doppler = 1`, which maps
v3's constant phase onto the per-path Doppler shift of the time-axis kernels (f_D * 1 s == -f_c (v tau / c + a tau^2 / 2c))."""
import os

import numpy as np
import pytest

from util import TOL_REL_FRO, per_user_rel_fro

HERE = os.path.dirname(os.path.abspath(__file__))


def _case():
    from make_golden_doppler_v3 import LIGHTSPEED, doppler_case
    d, cfg = doppler_case()
    g = np.load(os.path.join(HERE, "golden", "doppler_v3.npz"))
    return d, cfg, g, LIGHTSPEED


def _oracle_kw(cfg):
    return dict(bs_shape=cfg["bs_shape"], ue_shape=cfg["ue_shape"], bs_rotation=cfg["bs_rot"], ue_rotation=cfg["ue_rot"],
                subcarriers=cfg["n_sc"], selected_subcarriers=cfg["sel"], bandwidth=cfg["bandwidth"])


def test_oracle_reproduces_v3_doppler():
    from oracle import channel_oracle as orc
    d, cfg, g, c0 = _case()
    o0 = orc.compute_channels(d, **_oracle_kw(cfg))
    assert np.array_equal(o0["H"], g["H_static"])                   # v3 == v4 == oracle bit for bit without Doppler
    fd = orc.v3_doppler_shift(d["delay"], d["doppler_vel"], d["doppler_acc"], cfg["carrier_hz"])
    o1 = orc.compute_channels(d, **_oracle_kw(cfg), doppler_hz=fd, times=np.array([1.0]))
    err = per_user_rel_fro(o1["H"][..., 0], g["H_doppler"])
    assert err.max() <= 1e-6, err.max()
    # the golden is sensitive: ignoring Doppler misses it by two orders of magnitude more than the parity bar
    assert per_user_rel_fro(g["H_static"], g["H_doppler"]).max() > 100 * TOL_REL_FRO


def test_enable_doppler_host_mapping():
    import deepmimo_b200 as dmb
    from deepmimo_b200.channels import constant_doppler_shift
    from oracle import channel_oracle as orc
    d, cfg, g, c0 = _case()
    p = dmb.ChannelGenParameters()
    ds = dmb.Dataset(dict(d))
    assert constant_doppler_shift(ds, p) is None                    # flag off (the reference default, channel.py:50)
    p.enable_doppler = 1
    with pytest.raises(ValueError):
        constant_doppler_shift(ds, p)                               # no carrier frequency anywhere
    fd = constant_doppler_shift(ds, p, cfg["carrier_hz"])
    ref = orc.v3_doppler_shift(d["delay"], d["doppler_vel"], d["doppler_acc"], cfg["carrier_hz"])
    assert fd.dtype == np.float32 and np.array_equal(fd, ref, equal_nan=True)
    ds["rt_params"] = {"frequency": cfg["carrier_hz"]}
    assert np.array_equal(constant_doppler_shift(ds, p), fd, equal_nan=True)
    p.freq_domain = 0
    assert constant_doppler_shift(ds, p) is None                    # v3 applies no Doppler in the time-domain branch
    p.freq_domain = 1
    assert constant_doppler_shift(dmb.Dataset({k: v for k, v in d.items() if not k.startswith("doppler")}), p) is None
    p.ofdm.rx_filter = 1
    with pytest.raises(NotImplementedError):
        constant_doppler_shift(ds, p)


@pytest.mark.gpu
def test_gpu_enable_doppler_matches_v3_golden():
    import deepmimo_b200 as dmb
    d, cfg, g, c0 = _case()
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape, p.bs_antenna.rotation = cfg["bs_shape"], cfg["bs_rot"]
    p.ue_antenna.shape, p.ue_antenna.rotation = cfg["ue_shape"], cfg["ue_rot"]
    p.ofdm.subcarriers, p.ofdm.selected_subcarriers, p.ofdm.bandwidth = cfg["n_sc"], cfg["sel"], cfg["bandwidth"]
    H0 = dmb.Dataset(dict(d)).compute_channels(p, warn=False)
    assert H0.shape == g["H_static"].shape
    assert per_user_rel_fro(H0, g["H_static"]).max() <= TOL_REL_FRO
    p.enable_doppler = 1
    H1 = dmb.Dataset(dict(d)).compute_channels(p, warn=False, carrier_freq=cfg["carrier_hz"])
    assert H1.shape == g["H_doppler"].shape and H1.dtype == np.complex64 and H1.flags.c_contiguous
    err = per_user_rel_fro(H1, g["H_doppler"])
    assert err.max() <= TOL_REL_FRO, err.max()
    assert per_user_rel_fro(H1, g["H_static"]).max() > 10 * TOL_REL_FRO
    Ht = dmb.Dataset(dict(d)).compute_channels(p, warn=False, carrier_freq=cfg["carrier_hz"], out="torch")
    assert tuple(Ht.shape) == H1.shape and np.array_equal(Ht.cpu().numpy(), H1)
