"""fd_mma_kernel (warp-level tensor-core accumulate for small per-user outputs, dmk_fd_mma.cuh) against the oracle over random
shapes.  Cases cover: both chunk widths, K not a multiple of the chunk width (even and odd K, K = 1), panels up to 256 elements (J = 16 / 32 subcarriers, forced through DMK_WS_HELPERS), several groups of m-tiles
(DMK_WS_SPLIT = m-tiles resident at a time), partial m-tiles (chunks per user not a multiple of 16), a non-power-of-two number of
chunks per antenna row, FoV masks (every column runs its chain), dipole patterns, NaN holes, num_paths < n_cols, per-user UE
rotation, strided selections with an offset, 32 dense path columns (one user per pass), user counts that leave partial windows."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset, oracle_kwargs_from_params

pytestmark = pytest.mark.gpu

CASES = [
    # bs, ue, N, selection, n_users, fov, patterns, num_paths, holes, per-user rotation, n_cols, dense
    ((8, 1), (1, 1), 64, np.arange(64), 333, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((4, 2), (2, 1), 512, np.arange(128), 97, None, ("isotropic", "isotropic"), 25, True, True, 25, False),
    ((4, 4), (1, 1), 1024, np.arange(1024), 61, ((140, 120), (90, 80)), ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((2, 2), (2, 2), 2048, 5 + 3 * np.arange(304), 75, None, ("halfwave-dipole", "isotropic"), 10, True, True, 25, False),
    ((3, 1), (1, 1), 64, np.arange(16), 40, ((180, 90), (360, 180)), ("halfwave-dipole", "halfwave-dipole"), 25, True, False, 25, False),
    ((1, 1), (1, 1), 512, np.arange(16), 50, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((5, 1), (1, 3), 4096, 2 + 5 * np.arange(48), 129, None, ("isotropic", "isotropic"), 5, False, True, 25, False),
    ((8, 8), (1, 1), 64, np.arange(64), 150, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((8, 4), (1, 1), 256, np.arange(256), 77, ((120, 90), (180, 120)), ("isotropic", "halfwave-dipole"), 25, True, True, 25, False),
    ((8, 1), (1, 1), 64, np.arange(64), 90, None, ("isotropic", "isotropic"), 32, False, False, 32, True),
    ((4, 4), (2, 2), 128, np.arange(96), 45, None, ("isotropic", "isotropic"), 25, True, False, 25, True),
    ((16, 8), (1, 1), 64, np.arange(64), 70, None, ("isotropic", "isotropic"), 25, False, False, 25, False),                     # M = 128
    ((8, 8), (2, 2), 128, 3 + np.arange(64), 41, ((150, 100), (180, 120)), ("isotropic", "isotropic"), 25, True, True, 25, False),   # M = 256
    # K not a multiple of the chunk width: the last chunk of every antenna row is cut off (even K: 16-byte stores, odd K: 8-byte)
    ((8, 8), (1, 1), 128, np.arange(72), 60, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((4, 2), (2, 1), 512, np.arange(130), 97, None, ("isotropic", "isotropic"), 25, True, True, 25, False),
    ((3, 1), (1, 1), 64, np.arange(7), 40, ((180, 90), (360, 180)), ("halfwave-dipole", "halfwave-dipole"), 25, True, False, 25, False),
    ((2, 2), (2, 2), 2048, 5 + 3 * np.arange(301), 75, None, ("halfwave-dipole", "isotropic"), 10, True, True, 25, False),
    ((1, 1), (1, 1), 512, np.arange(1), 50, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    ((16, 1), (1, 1), 1024, np.arange(1000), 33, None, ("isotropic", "isotropic"), 25, False, False, 25, False),
    # panels of more than 256 elements (16-bit coordinate table) with few subcarriers
    ((32, 8), (2, 2), 512, np.arange(16), 21, None, ("isotropic", "isotropic"), 25, False, True, 25, False),                       # M = 1024
    ((16, 16), (2, 1), 512, 4 + 2 * np.arange(32), 19, ((150, 100), (180, 120)), ("isotropic", "isotropic"), 25, True, False, 25, False),   # M = 512
    ((32, 8), (2, 2), 64, np.arange(48), 9, None, ("halfwave-dipole", "isotropic"), 25, True, False, 25, False),
]


@pytest.mark.parametrize("variant", ["", "16", "32", "16:3"])      # chunk width by shape / forced; "16:3": three m-tiles per group
@pytest.mark.parametrize("case", range(len(CASES)))
def test_mma_kernel_matches_oracle(case, variant, monkeypatch):
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    bs, ue, n_sc, sel, n, fov, pats, num_paths, holes, per_user, n_cols, dense = CASES[case]
    monkeypatch.setenv("DMK_FD_KERNEL", "mma")
    j, _, grp = variant.partition(":")
    if j:
        monkeypatch.setenv("DMK_WS_HELPERS", j)
    if grp:
        monkeypatch.setenv("DMK_WS_SPLIT", grp)
    d = make_paths(n, 1900 + case, n_sc=n_sc, bandwidth=50e6, zero_frac=0.15, clip_frac=0.02, n_cols=n_cols, dense=dense)
    if holes:
        hole = np.random.default_rng(case).random(d["power"].shape) < 0.2
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].copy()
            d[k][hole] = np.nan
    ue_rot = np.random.default_rng(50 + case).uniform(-60, 60, (n, 3)) if per_user else np.array([10, -20, 30])
    p = {"bs_antenna": {"shape": np.array(bs), "spacing": 0.5, "rotation": np.array([5, 10, 20]), "radiation_pattern": pats[0]},
         "ue_antenna": {"shape": np.array(ue), "spacing": 0.4, "rotation": ue_rot, "radiation_pattern": pats[1]},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": num_paths, "freq_domain": 1,
         "ofdm": {"subcarriers": n_sc, "selected_subcarriers": sel, "bandwidth": 50e6, "rx_filter": 0}}
    bs_fov, ue_fov = (None, None) if fov is None else (np.array(fov[0]), np.array(fov[1]))
    H, info = make_dataset(dmb, d, bs_fov, ue_fov).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
    o = orc.compute_channels(d, **oracle_kwargs_from_params(p, bs_fov, ue_fov))
    assert info.kernel.startswith("fd_mma_kernel"), info.kernel
    if j == "16" or (j == "32" and len(sel) % 32 == 0):
        assert f"J={j}" in info.kernel, info.kernel
    err = assert_channels_close(H, o["H"], what=f"mma case {case} {variant!r}")
    assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.clip, o["clip"])
    if o["fov_mask"] is None:
        assert info.fov_mask is None
    else:
        assert np.array_equal(info.fov_mask, o["fov_mask"])
    print(f"case {case} {variant!r}: {info.kernel} max rel. Frobenius {err:.2e}")


def test_mma_kernel_handles_a_hundred_db_of_power_range():
    """The FP16 operands are scaled by the user's strongest path: a user whose paths span 100 dB, and users whose strongest path is
    very weak or very strong, stay inside the bar (relative to the user's own norm)."""
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    import os
    n = 64
    d = make_paths(n, 77, n_sc=64, bandwidth=50e6, zero_frac=0.0)
    rng = np.random.default_rng(3)
    pw = d["power"].copy()
    pw += rng.choice([-150.0, -60.0, 0.0, 80.0], size=(n, 1)).astype(np.float32)       # whole users shifted: -310 ... +20 dBW
    d["power"] = pw
    p = {"bs_antenna": {"shape": np.array([8, 1]), "spacing": 0.5, "rotation": np.array([0, 0, 0]), "radiation_pattern": "isotropic"},
         "ue_antenna": {"shape": np.array([1, 1]), "spacing": 0.5, "rotation": np.array([0, 0, 0]), "radiation_pattern": "isotropic"},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": 25, "freq_domain": 1,
         "ofdm": {"subcarriers": 64, "selected_subcarriers": np.arange(64), "bandwidth": 50e6, "rx_filter": 0}}
    os.environ["DMK_FD_KERNEL"] = "mma"
    try:
        H, info = make_dataset(dmb, d).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
    finally:
        os.environ.pop("DMK_FD_KERNEL")
    assert info.kernel.startswith("fd_mma_kernel"), info.kernel
    o = orc.compute_channels(d, **oracle_kwargs_from_params(p))
    assert_channels_close(H, o["H"], what="power range")


@pytest.mark.parametrize("bs_rot,ue_shape", [((0, 0, 0), (1, 1)), ((0, 0, 25), (1, 1)), ((5, 0, 0), (1, 1)), ((0, 0, 0), (2, 1))])
def test_single_element_unrotated_side_takes_the_short_chain(bs_rot, ue_shape, monkeypatch):
    """The reference's default UE (one element, rotation [0, 0, 0], isotropic, no FoV): its rotated angles enter the result only
    through their NaN-ness, and the kernel skips that side's float64 chain (dmk_prologue.cuh: side_angles_trivial).  Paths whose
    arrival angles are NaN / Inf while their power is finite must still drop out exactly as in the reference; a rotation about z
    alone keeps the short chain, a rotation about x or a second element does not (same results either way)."""
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    n = 300
    d = make_paths(n, 4242, n_sc=64, bandwidth=10e6, zero_frac=0.1)
    rng = np.random.default_rng(9)
    for key, val in (("aoa_az", np.nan), ("aoa_el", np.inf), ("aod_el", np.nan)):
        a = d[key].copy()
        hit = (rng.random(a.shape) < 0.05) & ~np.isnan(d["power"])
        a[hit] = val
        d[key] = a
    p = {"bs_antenna": {"shape": np.array([8, 1]), "spacing": 0.5, "rotation": np.array(bs_rot), "radiation_pattern": "isotropic"},
         "ue_antenna": {"shape": np.array(ue_shape), "spacing": 0.5, "rotation": np.array([0, 0, 0]), "radiation_pattern": "isotropic"},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": 25, "freq_domain": 1,
         "ofdm": {"subcarriers": 64, "selected_subcarriers": np.arange(64), "bandwidth": 10e6, "rx_filter": 0}}
    monkeypatch.setenv("DMK_FD_KERNEL", "mma")
    with np.errstate(invalid="ignore"):
        o = orc.compute_channels(d, **oracle_kwargs_from_params(p))
    H, info = make_dataset(dmb, d).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
    assert info.kernel.startswith("fd_mma_kernel"), info.kernel
    assert_channels_close(H, o["H"], what=f"trivial side {bs_rot} {ue_shape}")
    assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.clip, o["clip"])


def test_concurrent_launches_on_two_streams_draw_from_separate_counters():
    """fd_mma_kernel distributes users through a device counter that the last warp of a launch resets; launches in flight use
    different counter slots.  Two plans launched back to back on two streams (they overlap on the device), twenty rounds without a
    host synchronisation in between, must each produce what they produce alone."""
    import torch
    import deepmimo_b200 as dmb
    from deepmimo_b200 import _lib
    from deepmimo_b200.synth import scenario
    plans, refs = [], []
    for cfg, n in ((1, 30000), (6, 6000)):
        s = scenario(cfg, n)
        plan, _ = dmb.make_plan(make_dataset(dmb, s), dmb.ChannelGenParameters(s.params), warn=False)
        ref = plan.run(plan.alloc_out()).clone()
        assert _lib.last_kernel().startswith("fd_mma_kernel"), _lib.last_kernel()
        plans.append(plan); refs.append(ref)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [p.alloc_out() for p in plans]
    for rnd in range(20):
        for p, o, st in zip(plans, outs, streams):
            if rnd % 5 == 0:
                with torch.cuda.stream(st):
                    o.fill_(complex(float("nan"), 0.0))
            p.run(o, stream=st)
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o.view(torch.float32), r.view(torch.float32))
