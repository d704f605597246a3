"""Kernel choice across array shapes: time the FD path for a few (BS panel, UE panel, K) shapes with each eligible kernel.
    python tools/shape_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
shapes = [((32, 1), (1, 1), 512), ((16, 1), (1, 1), 1024), ((8, 4), (1, 1), 256), ((8, 8), (1, 1), 64), ((8, 8), (1, 1), 512),
          ((16, 8), (1, 1), 128), ((4, 4), (2, 1), 512), ((8, 2), (1, 1), 64), ((32, 4), (1, 1), 1024)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for bs, ue, k in shapes:
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(200000, (6 << 30) // (8 * m * k)))
    d = make_paths(n, 7, n_sc=max(k, 64), bandwidth=50e6, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.ofdm.subcarriers = max(k, 64); p.ofdm.selected_subcarriers = np.arange(k); p.ofdm.bandwidth = 50e6
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    row = []
    for var in ("auto", "tc", "ffma", "small"):
        if var == "auto": os.environ.pop("DMK_FD_KERNEL", None)
        else: os.environ["DMK_FD_KERNEL"] = var
        for _ in range(2): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        row.append(f"{var}:{_lib.last_kernel().split('<')[0][3:]} {ms:.3f} ms {8e-9 * n * m * k / (ms * 1e-3):.0f} GB/s")
    print(f"bs{bs} ue{ue} K={k} n={n} (M={m}): " + " | ".join(row), flush=True)
