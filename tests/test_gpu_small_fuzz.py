"""Small-array kernels (M = M_rx * M_tx <= 16) against the oracle over random shapes: the densely packed fd_small2_kernel (default)
and the round-1 one-warp-per-user kernel (DMK_FD_KERNEL=small1).  Cases cover the branches of the dense packing: FoV masks (every
column runs its chain), dipole patterns (float64 power), NaN holes inside the rows, num_paths < n_cols, per-user UE rotation with
several UE elements, two- and three-level delay-phasor seeds (K <= 64, <= 256, > 256), odd K, strided selections with an offset,
user counts that leave partial windows and ranges."""
import numpy as np
import pytest

from util import assert_channels_close, make_dataset, oracle_kwargs_from_params

pytestmark = pytest.mark.gpu

CASES = [
    # bs, ue, N, selection, n_users, fov, patterns, num_paths, holes, per-user rotation
    ((8, 1), (1, 1), 64, np.arange(64), 333, None, ("isotropic", "isotropic"), 25, False, False),
    ((4, 2), (2, 1), 512, np.arange(130), 97, None, ("isotropic", "isotropic"), 25, True, True),
    ((4, 4), (1, 1), 1024, np.arange(1024), 61, ((140, 120), (90, 80)), ("isotropic", "isotropic"), 25, False, False),
    ((2, 2), (2, 2), 2048, 5 + 3 * np.arange(300), 75, None, ("halfwave-dipole", "isotropic"), 10, True, True),
    ((3, 1), (1, 1), 64, np.arange(7), 40, ((180, 90), (360, 180)), ("halfwave-dipole", "halfwave-dipole"), 25, True, False),
    ((1, 1), (1, 1), 512, np.arange(1), 50, None, ("isotropic", "isotropic"), 25, False, False),
    ((5, 1), (1, 3), 4096, 2 + 5 * np.arange(33), 129, None, ("isotropic", "isotropic"), 5, False, True),
    ((16, 1), (1, 1), 4096, np.arange(4096), 19, None, ("isotropic", "isotropic"), 25, False, False),
    ((2, 4), (2, 1), 256, np.arange(255), 200, ((120, 90), (180, 120)), ("isotropic", "halfwave-dipole"), 25, True, True),
]


@pytest.mark.parametrize("variant", ["small", "small1"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_small_kernels_match_oracle(case, variant, monkeypatch):
    import deepmimo_b200 as dmb
    from deepmimo_b200 import _lib
    from deepmimo_b200.synth import make_paths
    from oracle import channel_oracle as orc
    bs, ue, n_sc, sel, n, fov, pats, num_paths, holes, per_user = CASES[case]
    monkeypatch.setenv("DMK_FD_KERNEL", variant)      # "small": the dense kernel even where the tensor-core kernel would take the shape
    d = make_paths(n, 900 + case, n_sc=n_sc, bandwidth=50e6, zero_frac=0.15, clip_frac=0.02)
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
    assert info.kernel.startswith("fd_small2_kernel" if variant == "small" else "fd_small_kernel<"), info.kernel
    err = assert_channels_close(H, o["H"], what=f"small case {case} {variant}")
    assert np.array_equal(info.valid, o["valid"]) and np.array_equal(info.clip, o["clip"])
    if o["fov_mask"] is None:
        assert info.fov_mask is None
    else:
        assert np.array_equal(info.fov_mask, o["fov_mask"])
    print(f"case {case} {variant}: {info.kernel.split(' ')[0]} max rel. Frobenius {err:.2e}")


def test_small2_is_deterministic_and_chunk_invariant():
    import deepmimo_b200 as dmb
    from deepmimo_b200.synth import scenario
    s = scenario(1, 5000)
    p = dmb.ChannelGenParameters(s.params)
    H1 = make_dataset(dmb, s).compute_channels(p, warn=False)
    H2 = make_dataset(dmb, s).compute_channels(p, warn=False, chunk_users=777)      # different warp ranges / windows, same results
    H3 = make_dataset(dmb, s).compute_channels(p, warn=False, out="torch").cpu().numpy()
    assert np.array_equal(H1, H2) and np.array_equal(H1, H3)
