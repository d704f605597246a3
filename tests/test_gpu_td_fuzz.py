"""td_warp_kernel (the reference's own time-domain mode: freq_domain = 0, no time axis; one warp per user, dmk_td.cuh) against the
oracle over random shapes: NaN holes inside the rows (the valid paths are compacted to the leading slots, channel.py:274-287), FoV
masks (masked paths keep their slot with a zero), dipole patterns, num_paths < n_cols, per-user UE rotation, several RX elements,
single-element panels, 32 path columns, float64 path matrices; masks and slot indices bit for bit.  DMK_FD_KERNEL=tile keeps the
CTA-per-user td_kernel (the kernel of the time-axis extension): both must agree."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset, oracle_kwargs_from_params

pytestmark = pytest.mark.gpu

CASES = [
    # bs, ue, n_users, fov, patterns, num_paths, holes, per-user rotation, n_cols, dense, float64
    ((8, 1), (1, 1), 333, None, ("isotropic", "isotropic"), 25, False, False, 25, False, False),
    ((4, 2), (2, 1), 97, None, ("isotropic", "isotropic"), 25, True, True, 25, False, False),
    ((8, 8), (1, 1), 61, ((140, 120), (90, 80)), ("isotropic", "isotropic"), 25, True, False, 25, False, False),
    ((2, 2), (2, 2), 75, None, ("halfwave-dipole", "isotropic"), 10, True, True, 25, False, False),
    ((1, 1), (1, 1), 50, ((180, 90), (360, 180)), ("halfwave-dipole", "halfwave-dipole"), 25, True, False, 25, False, False),
    ((5, 3), (1, 3), 129, None, ("isotropic", "isotropic"), 5, False, True, 25, False, False),
    ((16, 4), (2, 2), 40, None, ("isotropic", "isotropic"), 32, True, False, 32, True, False),
    ((8, 4), (2, 1), 200, ((120, 90), (180, 120)), ("isotropic", "halfwave-dipole"), 25, True, True, 25, False, True),
    ((3, 1), (1, 1), 7, None, ("isotropic", "isotropic"), 1, False, False, 1, False, False),
]


@pytest.mark.parametrize("variant", ["", "tile"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_td_kernels_match_oracle(case, variant, monkeypatch):
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    bs, ue, n, fov, pats, num_paths, holes, per_user, n_cols, dense, f64 = CASES[case]
    if variant:
        monkeypatch.setenv("DMK_FD_KERNEL", variant)
    d = make_paths(n, 2900 + case, n_sc=512, bandwidth=10e6, zero_frac=0.15, n_cols=n_cols, dense=dense)
    if holes:
        hole = np.random.default_rng(case).random(d["power"].shape) < 0.25
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].copy()
            d[k][hole] = np.nan
    if f64:
        rng = np.random.default_rng(77 + case)
        for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[k] = d[k].astype(np.float64) * (1.0 + 1e-9 * rng.standard_normal(d[k].shape))      # not float32-representable
    ue_rot = np.random.default_rng(50 + case).uniform(-60, 60, (n, 3)) if per_user else (np.array([0, 0, 0]) if case == 0 else np.array([10, -20, 30]))      # case 0: the default UE (short chain)
    p = {"bs_antenna": {"shape": np.array(bs), "spacing": 0.5, "rotation": np.array([5, 10, 20]), "radiation_pattern": pats[0]},
         "ue_antenna": {"shape": np.array(ue), "spacing": 0.4, "rotation": ue_rot, "radiation_pattern": pats[1]},
         "enable_doppler": 0, "enable_dual_polar": 0, "num_paths": num_paths, "freq_domain": 0,
         "ofdm": {"subcarriers": 512, "selected_subcarriers": np.arange(1), "bandwidth": 10e6, "rx_filter": 0}}
    bs_fov, ue_fov = (None, None) if fov is None else (np.array(fov[0]), np.array(fov[1]))
    H, info = make_dataset(dmb, d, bs_fov, ue_fov).compute_channels(dmb.ChannelGenParameters(p), return_info=True, warn=False)
    o = orc.compute_channels(d, **oracle_kwargs_from_params(p, bs_fov, ue_fov))
    assert info.kernel.startswith("td_kernel" if variant else "td_warp_kernel"), info.kernel
    err = assert_channels_close(H, o["H"], what=f"td case {case} {variant!r}")
    assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.path_slot, o["path_slot"])
    if o["fov_mask"] is None:
        assert info.fov_mask is None
    else:
        assert np.array_equal(info.fov_mask, o["fov_mask"])
    print(f"td case {case} {variant!r}: {info.kernel} max rel. Frobenius {err:.2e}")
