"""Where chained launches lose against one big launch on the city-scale shape (fd_ws_kernel, 512 KB per user), and what the work-item
granularity (DMK_WS_SPLIT: items per user) does to the tail:  python tools/launch_size.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepmimo_b200 as dmb
from deepmimo_b200.channels import chunk_is_independent
from deepmimo_b200.synth import scenario
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
N = 65536
s = scenario(5, N)
plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
big = plan.alloc_out(N)
ring = [plan.alloc_out(16384) for _ in range(3)]


def timed(step, reps=5):
    for _ in range(2): step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1); a.record(); step(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]


def report(name, ms, users=N):
    print(f"{name:70s} {ms:.3f} ms  {users * 512 * 1024 / ms * 1e-6:.0f} GB/s", flush=True)


for split in ("0", "2", "4"):
    os.environ["DMK_WS_SPLIT"] = split
    print(f"--- DMK_WS_SPLIT={split}")
    report("one launch, 65536 users, one buffer", timed(lambda: plan.run(big, 0, N)))
    for chunk in (16384, 8192):
        def step_ring():
            for i, a in enumerate(range(0, N, chunk)):
                plan.run(ring[i % 3][:chunk], a, a + chunk, independent=chunk_is_independent(i, 3))
        report(f"{N // chunk} launches of {chunk} into a ring of 3", timed(step_ring))
    for n1 in (8192, 2048):
        report(f"one launch of {n1} users alone", timed(lambda: plan.run(big[:n1], 0, n1)), n1)
