"""Per-launch device time inside a chain of fd_ws_kernel launches (events between the launches): ring of 3 against distinct slices of
one buffer, city-scale shape.    python tools/chain_gaps.py [users per launch] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepmimo_b200 as dmb
from deepmimo_b200.channels import chunk_is_independent
from deepmimo_b200.synth import scenario
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = scenario(5, L * n)
plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
big = plan.alloc_out(L * n)
ring = [plan.alloc_out(n) for _ in range(3)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for mode in ("slices", "ring3", "ring3-same-users"):
    for rep in range(3):
        flush.fill_(1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
        ev[0].record()
        for i in range(L):
            dst = big[i * n:(i + 1) * n] if mode == "slices" else ring[i % 3]
            lo = 0 if mode == "ring3-same-users" else i * n
            plan.run(dst, lo, lo + n, independent=chunk_is_independent(i, 3))
            ev[i + 1].record()
        torch.cuda.synchronize()
    d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(L)]
    print(f"{mode:18s} total {sum(d):7.0f} us | per launch: " + " ".join(f"{x:5.0f}" for x in d), flush=True)
