"""fd_rows_kernel (a handful of selected subcarriers, K <= 8 -- the reference's DEFAULT parameters select one; warp per user with
lanes = antenna rows, dmk_fd_rows.cuh) against the oracle: K = 1 .. 8, affine selections and arbitrary lists, panels from one element
to 32 x 8 x 2 x 2 (row blocks with idle lanes and many blocks), FoV masks, dipole patterns, NaN holes, num_paths < n_cols, per-user UE
rotation, 32 path columns, float64 path matrices; masks bit for bit."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset, oracle_kwargs_from_params

pytestmark = pytest.mark.gpu

CASES = [
    # bs, ue, N, selection, n_users, fov, patterns, num_paths, holes, per-user rotation, n_cols, dense, float64
    ((8, 1), (1, 1), 512, np.arange(1), 333, None, ("isotropic", "isotropic"), 25, False, False, 25, False, False),       # the reference's default call
    ((8, 8), (1, 1), 512, np.array([0]), 97, None, ("isotropic", "isotropic"), 25, True, False, 25, False, False),
    ((32, 8), (2, 2), 512, np.array([5]), 23, None, ("isotropic", "isotropic"), 25, False, True, 25, False, False),        # M = 1024: 32 row blocks
    ((4, 2), (2, 1), 512, np.array([3, 17, 100]), 61, ((140, 120), (90, 80)), ("isotropic", "isotropic"), 25, True, True, 25, False, False),
    ((5, 3), (1, 3), 4096, 2 + 5 * np.arange(8), 129, None, ("halfwave-dipole", "isotropic"), 5, False, True, 25, False, False),
    ((1, 1), (1, 1), 64, np.arange(2), 50, ((180, 90), (360, 180)), ("halfwave-dipole", "halfwave-dipole"), 25, True, False, 25, False, False),
    ((16, 4), (1, 1), 1024, np.array([1023, 0, 512, 7, 7]), 40, None, ("isotropic", "isotropic"), 32, True, False, 32, True, False),
    ((8, 4), (2, 1), 256, np.arange(4), 200, ((120, 90), (180, 120)), ("isotropic", "halfwave-dipole"), 25, True, True, 25, False, True),
    ((3, 3), (2, 2), 128, np.arange(7), 75, None, ("isotropic", "isotropic"), 10, True, False, 25, False, False),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_rows_kernel_matches_oracle(case):
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    bs, ue, n_sc, sel, n, fov, pats, num_paths, holes, per_user, n_cols, dense, f64 = CASES[case]
    d = make_paths(n, 3900 + case, n_sc=n_sc, bandwidth=n_sc / 4.2e-6 if n_sc < 512 else 50e6, zero_frac=0.15, clip_frac=0.02, n_cols=n_cols, dense=dense)
    if holes:
        hole = np.random.default_rng(case).random(d["power"].shape) < 0.2
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].copy()
            d[k][hole] = np.nan
    if f64:
        rng = np.random.default_rng(77 + case)
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].astype(np.float64) * (1.0 + 1e-9 * rng.standard_normal(d[k].shape))
    ue_rot = np.random.default_rng(50 + case).uniform(-60, 60, (n, 3)) if per_user else (np.array([0, 0, 0]) if case in (0, 1) else np.array([10, -20, 30]))      # cases 0, 1: the default UE (short chain)
    p = {"bs_antenna": {"shape": np.array(bs), "spacing": 0.5, "rotation": np.array([5, 10, 20]), "radiation_pattern": pats[0]},
         "ue_antenna": {"shape": np.array(ue), "spacing": 0.4, "rotation": ue_rot, "radiation_pattern": pats[1]},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": num_paths, "freq_domain": 1,
         "ofdm": {"subcarriers": n_sc, "selected_subcarriers": sel, "bandwidth": n_sc / 4.2e-6 if n_sc < 512 else 50e6, "rx_filter": 0}}
    bs_fov, ue_fov = (None, None) if fov is None else (np.array(fov[0]), np.array(fov[1]))
    H, info = make_dataset(dmb, d, bs_fov, ue_fov).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
    o = orc.compute_channels(d, **oracle_kwargs_from_params(p, bs_fov, ue_fov))
    assert info.kernel.startswith("fd_rows_kernel"), info.kernel
    err = assert_channels_close(H, o["H"], what=f"rows case {case}")
    assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.clip, o["clip"])
    if o["fov_mask"] is None:
        assert info.fov_mask is None
    else:
        assert np.array_equal(info.fov_mask, o["fov_mask"])
    print(f"rows case {case}: {info.kernel} max rel. Frobenius {err:.2e}")
