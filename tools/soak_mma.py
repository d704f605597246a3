"""Soak test of fd_mma_kernel: random shapes it is eligible for (M <= 256, K <= 4096, mostly multiples of 16), both chunk widths, one to
three m-tiles per group, selections with offset / stride, FoV / dipole / holes / num_paths / per-user UE rotation, user counts and
chunkings (chunks are launched back to back: the ticket counter must be back at zero every time); every result compared with the
packed-FP32 CUDA-core kernel and the generic tile kernel, masks bit for bit.   python tools/soak_mma.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
from util import per_user_rel_fro
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 97531)
t0 = time.time(); n_cases = 0; worst = 0.0; kinds = {}
while time.time() - t0 < budget:
    while True:
        bs = (int(rng.integers(1, 17)), int(rng.integers(1, 17))); ue = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
        m = bs[0] * bs[1] * ue[0] * ue[1]
        if m <= 256: break
    k = 16 * int(rng.choice([1, 2, 3, 4, 5, 8, 12, 16, 24, 32, 64, 100, 256]))
    if rng.random() < 0.3: k = int(rng.choice([1, 2, 7, 52, 63, 65, 72, 130, 257, 1000]))      # not a multiple of the chunk width
    step = int(rng.choice([1, 1, 3])); start = int(rng.integers(0, 5))
    n_sc = int(2 ** np.ceil(np.log2(start + step * k + 1)))
    n = int(rng.choice([1, 3, 63, 64, 65, 500, 2049, 9000]))
    n = max(1, min(n, (1 << 27) // (8 * m * k)))
    n_cols = int(rng.choice([1, 7, 25, 32]))
    bw = float(rng.choice([n_sc / 4.2e-6, 3 * n_sc / 4.2e-6]))             # second choice: two thirds of the paths are clipped
    d = make_paths(n, int(rng.integers(1, 10 ** 6)), n_sc=n_sc, bandwidth=bw, n_cols=n_cols, zero_frac=float(rng.choice([0.0, 0.1, 0.9])),
                   clip_frac=0.02, dense=bool(rng.random() < 0.2))
    if rng.random() < 0.5:
        hole = rng.random(d["power"].shape) < 0.2
        for key in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"):
            d[key] = d[key].copy(); d[key][hole] = np.nan
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue)
    p.bs_antenna.rotation = np.array([5, 10, 15]) if rng.random() < 0.7 else np.array([0, 0, 40])
    r = rng.random()
    p.ue_antenna.rotation = rng.uniform(-60, 60, (n, 3)) if r < 0.3 else (np.array([0, 0, 0]) if r < 0.7 else np.array([10, -20, 30]))
    p.bs_antenna.radiation_pattern = str(rng.choice(["isotropic", "isotropic", "halfwave-dipole"]))
    p.num_paths = int(rng.choice([n_cols, max(1, n_cols // 2)]))
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = start + step * np.arange(k); p.ofdm.bandwidth = bw
    ds = dmb.Dataset(d)
    if rng.random() < 0.3:
        ds.apply_fov(bs_fov=np.array([140, 120]), ue_fov=np.array([180, 120]))
    plan, _ = dmb.make_plan(ds, p, warn=False)
    res = {}
    for var in ("ffma", "tile", "mma"):
        os.environ["DMK_FD_KERNEL"] = var
        os.environ.pop("DMK_WS_HELPERS", None); os.environ.pop("DMK_WS_SPLIT", None)
        if var == "mma":
            j = str(rng.choice(["", "16", "32"])); grp = str(rng.choice(["", "1", "2", "3"]))
            if j: os.environ["DMK_WS_HELPERS"] = j
            if grp: os.environ["DMK_WS_SPLIT"] = grp
        out, masks = plan.alloc_out(), plan.alloc_masks()
        out.fill_(complex(float("nan"), 0.0))
        chunk = n if var != "mma" else max(1, n // int(rng.integers(1, 5)))
        for a in range(0, n, chunk):
            b = min(a + chunk, n)
            plan.run(out[a:b], a, b, {kk: v[a:b] for kk, v in masks.items()})
        if var == "mma":
            kern = _lib.last_kernel()
            assert kern.startswith("fd_mma_kernel"), kern
            key = kern.split(">")[0].split("3xf16,")[1]
            kinds[key] = kinds.get(key, 0) + 1
        res[var] = (out.cpu().numpy(), {kk: v.cpu().numpy() for kk, v in masks.items()})
    assert not np.isnan(res["mma"][0].view(np.float32)).any(), ("unwritten / NaN output", bs, ue, k, n, n_cols)
    for var in ("ffma", "tile"):
        err = per_user_rel_fro(res["mma"][0], res[var][0])
        e = float(err.max()) if err.size else 0.0
        worst = max(worst, e)
        assert e < 3e-6, (e, var, bs, ue, k, n, n_cols, kern)
        for kk in res[var][1]:
            assert np.array_equal(res["mma"][1][kk], res[var][1][kk]), (kk, var, bs, ue, k, n)
    n_cases += 1
for v in ("DMK_FD_KERNEL", "DMK_WS_HELPERS", "DMK_WS_SPLIT"): os.environ.pop(v, None)
print(f"mma soak OK: {n_cases} cases, worst per-user rel. Frobenius between kernels {worst:.2e}, {time.time() - t0:.0f} s; instantiations {kinds}")
