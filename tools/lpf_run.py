"""Receive-LPF (ofdm.rx_filter = 1) timing on a BASELINE shape:  python tools/lpf_run.py cfg2 1024"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
cfg, n = int(sys.argv[1][3:]), int(sys.argv[2])
s = scenario(cfg, n)
for lpf in (0, 1):
    p = dmb.ChannelGenParameters(s.params); p.ofdm.rx_filter = lpf
    ds = dmb.Dataset(dict(s.data))
    if s.bs_fov is not None: ds.apply_fov(bs_fov=s.bs_fov, ue_fov=s.ue_fov)
    plan, _ = dmb.make_plan(ds, p, warn=False)
    out = plan.alloc_out()
    for _ in range(2): plan.run(out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): plan.run(out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{s.name} n={n} rx_filter={lpf}: {ms:.3f} ms per pass, {out.numel() / ms / 1e6:.1f} Gcoef/s  {_lib.last_kernel()}")
if len(sys.argv) > 3:
    from oracle import channel_oracle as orc
    from util import oracle_kwargs_from_params
    m = int(sys.argv[3]); kw = oracle_kwargs_from_params(p, s.bs_fov, s.ue_fov)
    t0 = time.perf_counter(); orc.compute_channels(s.data, **kw, user_range=(0, m)); t1 = time.perf_counter()
    print(f"oracle (NumPy, 1 core) rx_filter=1: {m} users in {t1 - t0:.2f} s = {m * out[0].numel() / (t1 - t0) / 1e6:.2f} Mcoef/s")
