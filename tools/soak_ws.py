"""Soak test of fd_ws_kernel: many launches of random shapes and user counts (both helper configurations, independent-launch chains),
every result compared with the packed-FP32 kernel.  python tools/soak_ws.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
from util import per_user_rel_fro
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(12345)
t0 = time.time(); n_cases = 0; n_launch = 0; worst = 0.0
while time.time() - t0 < budget:
    while True:
        bs = (int(rng.integers(1, 33)), int(rng.integers(1, 17))); ue = (int(rng.integers(1, 3)), int(rng.integers(1, 3)))
        m = bs[0] * bs[1] * ue[0] * ue[1]
        if 16 <= m <= 1024: break
    k = 64 * int(rng.integers(1, 17)); step = int(rng.choice([1, 2])); start = int(rng.integers(0, 3))
    n_sc = int(2 ** np.ceil(np.log2(start + step * k + 1)))
    n = int(rng.choice([1, 2, 37, 295, 296, 297, 1000, 5000])); n = max(1, min(n, (1 << 29) // (8 * m * k)))
    d = make_paths(n, int(rng.integers(1, 10 ** 6)), n_sc=n_sc, bandwidth=50e6, n_cols=int(rng.choice([3, 25, 32])), zero_frac=float(rng.choice([0.0, 0.1, 0.9])))
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.num_paths = d["power"].shape[1]
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = start + step * np.arange(k); p.ofdm.bandwidth = 50e6
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    os.environ["DMK_FD_KERNEL"] = "ffma"
    ref = plan.run(plan.alloc_out()).cpu().numpy()
    os.environ["DMK_FD_KERNEL"] = "tc"
    os.environ["DMK_WS_HELPERS"] = str(rng.choice([1, 4]))
    chunk = max(1, n // int(rng.integers(1, 6)))
    outs = []
    for rep in range(3):                                         # back-to-back chains of independent launches
        bufs = [plan.alloc_out(min(chunk, n - a)) for a in range(0, n, chunk)]
        for i, a in enumerate(range(0, n, chunk)):
            plan.run(bufs[i], a, min(a + chunk, n), independent=(i > 0)); n_launch += 1
        outs.append(bufs)
    torch.cuda.synchronize()
    for bufs in outs:
        got = np.concatenate([b.cpu().numpy() for b in bufs], axis=0)
        err = per_user_rel_fro(got, ref)
        e = float(err.max()) if err.size else 0.0
        worst = max(worst, e)
        assert e < 2e-6, (e, bs, ue, k, n, chunk, _lib.last_kernel())
    n_cases += 1
print(f"soak OK: {n_cases} cases, {n_launch} fd_ws launches, worst per-user rel. Frobenius vs FP32 kernel {worst:.2e}, {time.time() - t0:.0f} s")
