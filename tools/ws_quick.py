"""Quick timing of the persistent tensor-core kernel on the headline shapes (cfg2 4096 users, cfg5 65536 users in one launch, dense
variants), CUDA events, L2 flushed:  python tools/ws_quick.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for cfg, n, dense in ((2, 4096, False), (5, 65536, False), (2, 4096, True), (5, 32768, True)):
    s = scenario(cfg, n, dense=dense)
    plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
    out = plan.alloc_out()
    for _ in range(3): plan.run(out)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in ev:
        flush.fill_(1); a.record(); plan.run(out); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    print(f"cfg{cfg}{' dense' if dense else ''} n={n}: min {ms[0]:.3f} med {med:.3f} ms  {out.numel() * 8e-9 / (med * 1e-3):.0f} GB/s  [{_lib.last_kernel()}]", flush=True)
    del out, plan
    torch.cuda.empty_cache()
