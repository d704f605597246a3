"""Draw sizes of fd_mma_kernel's work distribution (DMK_WS_SPLIT = 100 + users per draw, 100 = guided 8 / 4 / 2):  python tools/chunk_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for cfg, n in ((1, 80000), (6, 131072)):
    s = scenario(cfg, n)
    plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
    out = plan.alloc_out()
    row = []
    for c in (0, 2, 4, 8):      # 0: guided draws (8 / 4 / 2)
        os.environ["DMK_WS_SPLIT"] = str(100 + c)
        for _ in range(3): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(9)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[4]
        row.append(f"chunk {c}: {ms:.4f} ms")
    print(s.name, " | ".join(row), flush=True)
