"""Time-domain kernel on small per-user outputs:  python tools/td_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for bs, ue in (((8, 1), (1, 1)), ((8, 8), (1, 1)), ((8, 4), (2, 1)), ((32, 8), (2, 2))):
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(400000, (4 << 30) // (8 * m * 25)))
    d = make_paths(n, 7, n_sc=512, bandwidth=10e6, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15]); p.freq_domain = 0
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    for _ in range(3): plan.run(out)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
    for a, b in ev:
        flush.fill_(1); a.record(); plan.run(out); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[3]
    print(f"TD bs{bs} ue{ue} n={n} ({8 * m * 25} B/user): {ms:.3f} ms {out.numel() * 8e-9 / (ms * 1e-3):.0f} GB/s  {_lib.last_kernel()}", flush=True)
