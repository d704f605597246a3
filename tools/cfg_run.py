"""One BASELINE workload through the FD/TD path a few times (for ncu):  python tools/cfg_run.py cfg1 [users] [dense]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
cfg = int(sys.argv[1].replace("cfg", ""))
users = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
s = scenario(cfg, users, dense=(len(sys.argv) > 3 and sys.argv[3] == "dense"))
ds = dmb.Dataset(dict(s.data))
if s.bs_fov is not None:
    ds.apply_fov(bs_fov=s.bs_fov, ue_fov=s.ue_fov)
plan, _ = dmb.make_plan(ds, dmb.ChannelGenParameters(s.params), times=s.times, doppler=s.doppler_hz, warn=False)
out = plan.alloc_out()
for _ in range(4):
    plan.run(out)
torch.cuda.synchronize()
print(_lib.last_kernel())
