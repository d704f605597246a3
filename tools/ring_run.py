"""The bench's city-scale leg in isolation: 200 000 users of one base station through a ring of 8 192-user chunks, whole-step
CUDA events; variants to locate what a step loses against 200 000 x (single-launch time per user).   python tools/ring_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepmimo_b200 as dmb
from deepmimo_b200.channels import chunk_is_independent
from deepmimo_b200.synth import scenario
N = 200000
s = scenario(5, N)
plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(step, reps=4, do_flush=True):
    for _ in range(2): step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        if do_flush: flush.fill_(1)
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]


def chain(ring, chunk, users, rule):
    def step():
        for i, a in enumerate(range(0, users, chunk)):
            b = min(a + chunk, users)
            plan.run(ring[i % len(ring)][: b - a], a, b, independent=rule(i, len(ring)))
    return step


for chunk, n_ring in ((8192, 3), (8192, 6), (16384, 3), (32768, 3)):
    ring = [plan.alloc_out(chunk) for _ in range(n_ring)]
    for name, rule in (("ring rule", chunk_is_independent), ("all plain", lambda i, r: False)):
        ms = timed(chain(ring, chunk, N, rule))
        print(f"200k users, chunks of {chunk}, ring of {n_ring}, {name:9s}: {ms:.2f} ms  {N * 512 * 1024 / ms * 1e-6:.0f} GB/s", flush=True)
    if chunk == 8192 and n_ring == 3:
        ms = timed(chain(ring, chunk, 65536, chunk_is_independent))
        print(f"   first 65536 users only: {ms:.2f} ms  {65536 * 512 * 1024 / ms * 1e-6:.0f} GB/s", flush=True)
        ms = timed(chain(ring, chunk, N, chunk_is_independent), do_flush=False)
        print(f"   without the L2 flush between steps: {ms:.2f} ms", flush=True)
    del ring
    torch.cuda.empty_cache()
big = plan.alloc_out(131072)
ms = timed(lambda: plan.run(big, 0, 131072))
print(f"one launch of 131072 users: {ms:.2f} ms  {131072 * 512 * 1024 / ms * 1e-6:.0f} GB/s")
ms = timed(lambda: plan.run(big, 68928, 200000))
print(f"one launch of users 68928..200000: {ms:.2f} ms  {131072 * 512 * 1024 / ms * 1e-6:.0f} GB/s")
