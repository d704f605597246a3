"""Soak test of fd_small2_kernel: random small-array shapes (M <= 16), selections, FoV / dipole / holes / num_paths, user counts and
chunkings, every result compared with the round-1 one-warp-per-user kernel and the generic tile kernel.  python tools/soak_small.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
from util import per_user_rel_fro
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(4321)
t0 = time.time(); n_cases = 0; worst = 0.0
while time.time() - t0 < budget:
    while True:
        bs = (int(rng.integers(1, 17)), int(rng.integers(1, 5))); ue = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
        m = bs[0] * bs[1] * ue[0] * ue[1]
        if m <= 16: break
    k = int(rng.choice([1, 2, 7, 63, 64, 65, 130, 256, 257, 1000, 4096])); step = int(rng.choice([1, 3])); start = int(rng.integers(0, 5))
    n_sc = int(2 ** np.ceil(np.log2(start + step * k + 1)))
    n = int(rng.choice([1, 3, 63, 64, 65, 500, 4097]))
    n = max(1, min(n, (1 << 27) // (8 * m * k)))
    n_cols = int(rng.choice([1, 7, 25, 32]))
    d = make_paths(n, int(rng.integers(1, 10 ** 6)), n_sc=n_sc, bandwidth=50e6, n_cols=n_cols, zero_frac=float(rng.choice([0.0, 0.1, 0.9])), clip_frac=0.02)
    if rng.random() < 0.5:
        hole = rng.random(d["power"].shape) < 0.2
        for key in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[key] = d[key].copy(); d[key][hole] = np.nan
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.bs_antenna.radiation_pattern = str(rng.choice(["isotropic", "halfwave-dipole"]))
    p.num_paths = int(rng.choice([n_cols, max(1, n_cols // 2)]))
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = start + step * np.arange(k); p.ofdm.bandwidth = 50e6
    ds = dmb.Dataset(d)
    if rng.random() < 0.4:
        ds.apply_fov(bs_fov=np.array([140, 120]), ue_fov=np.array([180, 120]))
    plan, _ = dmb.make_plan(ds, p, warn=False)
    res = {}
    for var in ("small1", "tile", "small"):
        os.environ["DMK_FD_KERNEL"] = var
        out, masks = plan.alloc_out(), plan.alloc_masks()
        chunk = n if var != "small" else max(1, n // int(rng.integers(1, 4)))
        for a in range(0, n, chunk):
            b = min(a + chunk, n)
            plan.run(out[a:b], a, b, {kk: v[a:b] for kk, v in masks.items()})
        if var == "small":
            assert _lib.last_kernel().startswith("fd_small2_kernel"), _lib.last_kernel()
        res[var] = (out.cpu().numpy(), {kk: v.cpu().numpy() for kk, v in masks.items()})
    for var in ("small1", "tile"):
        err = per_user_rel_fro(res["small"][0], res[var][0])
        e = float(err.max()) if err.size else 0.0
        worst = max(worst, e)
        assert e < 3e-6, (e, var, bs, ue, k, n, n_cols)
        for kk in res[var][1]:
            assert np.array_equal(res["small"][1][kk], res[var][1][kk]), (kk, var, bs, ue, k, n)
    n_cases += 1
print(f"small soak OK: {n_cases} cases, worst per-user rel. Frobenius between kernels {worst:.2e}, {time.time() - t0:.0f} s")
