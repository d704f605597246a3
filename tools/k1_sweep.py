"""The reference's DEFAULT call (ofdm.selected_subcarriers = [0]: one subcarrier) and other tiny selections:  python tools/k1_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for bs, ue, k in (((8, 1), (1, 1), 1), ((8, 8), (1, 1), 1), ((32, 8), (2, 2), 1), ((8, 1), (1, 1), 4), ((8, 8), (1, 1), 8)):
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = 200000
    d = make_paths(n, 7, n_sc=512, bandwidth=10e6, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue)
    p.ofdm.selected_subcarriers = np.arange(k)
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    row = []
    for var in ("auto", "mma", "small", "ffma", "tile"):
        if var == "auto": os.environ.pop("DMK_FD_KERNEL", None)
        else: os.environ["DMK_FD_KERNEL"] = var
        for _ in range(3): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[3]
        row.append(f"{var}:{_lib.last_kernel().split('<')[0][3:]} {ms:.3f} ms")
    os.environ.pop("DMK_FD_KERNEL", None)
    byt = n * (8 * m * k + 7 * 4 * 25)
    print(f"bs{bs} ue{ue} K={k} n={n} (M={m}, {8 * m * k} B/user, HBM bound {byt / 6.5e12 * 1e3:.3f} ms): " + " | ".join(row), flush=True)
