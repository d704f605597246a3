"""Per-CTA timeline of three chained fd_ws_kernel launches (debug build: DMK_NVCC_EXTRA=-DDMK_TC_TRACE python -m deepmimo_b200.build
--force): kernel entry, first user picked up by the drain warps, exit, users per CTA -- where a chain of launches loses time
against the steady state of one big launch.     python tools/ws_trace.py [users per launch] [independent: 0|1]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
indep = len(sys.argv) > 2 and sys.argv[2] == "1"
s = scenario(5, 3 * n)
plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
ring = [plan.alloc_out(n) for _ in range(3)]
lib = _lib.load()
for rep in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(3):
        plan.run(ring[i], i * n, (i + 1) * n, independent=(indep and i > 0))
    b.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 4096)()
    lib.dmk_debug_tc_trace(buf, 4096)
    t = np.array(buf[:3552], dtype=np.int64).reshape(3, 296, 4)
    order = np.argsort(t[:, :, 0].min(axis=1))
    t0 = t[:, :, 0].min()
    print(f"rep {rep}: three launches of {n} users, event time {a.elapsed_time(b) * 1e3:.0f} us (independent={indep})")
    for k in order:
        e, f, x, u = (t[k, :, 0] - t0) / 1e3, (t[k, :, 1] - t0) / 1e3, (t[k, :, 2] - t0) / 1e3, t[k, :, 3]
        print(f"   launch: entry {e.min():7.1f}..{e.max():7.1f} (median {np.median(e):7.1f}) | first user median {np.median(f):7.1f} max {f.max():7.1f} | "
              f"exit {x.min():7.1f}..{x.max():7.1f} (median {np.median(x):7.1f}) us | users/CTA {u.min()}..{u.max()}")
