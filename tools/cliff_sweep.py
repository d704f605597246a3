"""Shapes between the kernel families (large panels with few subcarriers, arbitrary subcarrier lists):  python tools/cliff_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(0)
cases = [((32, 8), (2, 2), np.arange(16)), ((32, 8), (2, 2), np.arange(32)), ((32, 8), (2, 2), np.arange(48)), ((32, 8), (1, 1), np.arange(16)),
         ((16, 16), (2, 1), np.arange(32)), ((8, 8), (1, 1), np.sort(rng.choice(512, 64, replace=False))), ((32, 8), (2, 2), np.sort(rng.choice(512, 64, replace=False))),
         ((8, 8), (1, 1), np.arange(12))]
for bs, ue, sel in cases:
    m = bs[0] * bs[1] * ue[0] * ue[1]; k = len(sel)
    n = int(min(200000, (4 << 30) // (8 * m * k)))
    d = make_paths(n, 7, n_sc=512, bandwidth=50e6, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.ofdm.bandwidth = 50e6
    p.ofdm.selected_subcarriers = sel
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    for _ in range(3): plan.run(out)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        flush.fill_(1); a.record(); plan.run(out); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
    aff = bool(np.all(np.diff(sel) == (sel[1] - sel[0])))
    print(f"bs{bs} ue{ue} K={k}{'' if aff else ' (list)'} n={n} (M={m}, {8 * m * k // 1024} KB/user): {ms:.3f} ms {out.numel() * 8e-9 / (ms * 1e-3):.0f} GB/s  {_lib.last_kernel().split(' ')[0]}", flush=True)
