"""Fused beam amplitude map on a BASELINE shape (for ncu / timing):  python tools/beam_run.py cfg2 4096 16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200.synth import scenario
cfg, n, nb = int(sys.argv[1][3:]), int(sys.argv[2]), int(sys.argv[3])
s = scenario(cfg, n)
plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
F = np.array([dmb.steering_vec(list(plan.spec.bs_shape), phi=a).squeeze() for a in np.linspace(-60, 60, nb)]).astype(np.complex64)
Fd = torch.from_numpy(F).cuda()
amp = torch.empty((plan.n_users, nb), dtype=torch.float32, device="cuda")
for _ in range(3):
    plan.run_beams(Fd, amp)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    plan.run_beams(Fd, amp)
b.record(); torch.cuda.synchronize()
from deepmimo_b200 import _lib
print(_lib.last_kernel(), f"{a.elapsed_time(b) / 5:.3f} ms per call")
