"""One mid-size shape through the FD path (for ncu):  python tools/mid_run.py 8 8 64 100000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
b0, b1, k, n = (int(v) for v in sys.argv[1:5])
bw = min(50e6, max(k, 64) / 4.2e-6)
d = make_paths(n, 7, n_sc=max(k, 64), bandwidth=bw, n_cols=25)
p = dmb.ChannelGenParameters()
p.bs_antenna.shape = np.array([b0, b1]); p.bs_antenna.rotation = np.array([5, 10, 15])
p.ofdm.subcarriers = max(k, 64); p.ofdm.selected_subcarriers = np.arange(k); p.ofdm.bandwidth = bw
plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
out = plan.alloc_out()
for _ in range(4): plan.run(out)
torch.cuda.synchronize()
print(_lib.last_kernel())
