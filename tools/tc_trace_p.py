"""Phase timing of one mid-grid CTA of fd_tc_persist_kernel (debug build: DMK_NVCC_EXTRA=-DDMK_TC_TRACE python -m deepmimo_b200.build --force).
    python tools/tc_trace_p.py cfg5 20000"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import scenario
cfg, n = int(sys.argv[1][3:]), int(sys.argv[2])
s = scenario(cfg, n)
ds = dmb.Dataset(dict(s.data))
plan, _ = dmb.make_plan(ds, dmb.ChannelGenParameters(s.params), warn=False)
out = plan.alloc_out()
for _ in range(3):
    plan.run(out)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 4096)()
lib.dmk_debug_tc_trace(buf, 4096)
t = np.array(buf[:], dtype=np.int64)
print("kernel", _lib.last_kernel())
t0 = t[0]
print("stage | worker: wait build sync epilogue | issuer")
for st in range(0, 40):
    w = t[16 * st: 16 * st + 5]; u = t[16 * st + 8: 16 * st + 13]
    if w[0] == 0: break
    print(f"{st:3d} | top@{w[0]-t0:7d}  wait {w[1]-w[0]:6d}  build {w[2]-w[1]:6d}  sync {w[3]-w[2]:6d}  epi {w[4]-w[3]:6d} | "
          f"issuer: top@{u[0]-t0:7d} wait {u[1]-u[0]:6d} sync_done@{u[3]-t0:7d} issue {u[4]-u[3]:6d}")
print("per-user (6 consecutive users of the CTA): np | workers busy | helper busy | user period")
for i in range(6):
    b = 4000 + 8 * i
    if t[b] == 0: continue
    nxt = t[b + 8] if i < 5 and t[b + 8] else 0
    print(f"  np {t[b+5]:2d} | workers {t[b+1]-t[b]:7d} | helper {t[b+3]-t[b+2]:7d} | period {nxt - t[b] if nxt else -1:7d}")
