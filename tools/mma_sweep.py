"""fd_mma_kernel against the other kernel families on the small and the helper-bound shapes:  python tools/mma_sweep.py [quick]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths, scenario
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(plan, out, variants):
    row = []
    for var in variants:
        kern, _, rest = var.partition("/")
        j, _, grp = rest.partition(":")
        for k, v in (("DMK_FD_KERNEL", kern if kern != "auto" else ""), ("DMK_WS_HELPERS", j), ("DMK_WS_SPLIT", grp)):
            if v: os.environ[k] = v
            else: os.environ.pop(k, None)
        for _ in range(3): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
        row.append(f"{var}:{_lib.last_kernel().split('<')[0][3:]} {ms:.3f} ms {out.numel() * 8e-9 / (ms * 1e-3):.0f} GB/s")
    for k in ("DMK_FD_KERNEL", "DMK_WS_HELPERS", "DMK_WS_SPLIT"): os.environ.pop(k, None)
    return " | ".join(row)


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for dense in (False, True):
    s = scenario(1, 80000, dense=dense)
    plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
    out = plan.alloc_out()
    print(f"cfg1{' dense' if dense else ''} (8x1, K=64, 80k users): " + timed(plan, out, ("mma/16:1", "mma/16:2", "mma/32", "small")), flush=True)
    print("   ", _lib.last_kernel(), flush=True)
shapes = [((8, 2), (1, 1), 64), ((4, 4), (1, 1), 64), ((16, 1), (1, 1), 1024), ((4, 2), (2, 1), 256), ((8, 4), (1, 1), 64),
          ((8, 8), (1, 1), 64), ((8, 8), (1, 1), 128), ((8, 4), (1, 1), 256), ((8, 8), (1, 1), 512)]
for bs, ue, k in (shapes[:1] + shapes[5:6] if quick else shapes):
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(200000, (4 << 30) // (8 * m * k)))
    bw = min(50e6, max(k, 64) / 4.2e-6)      # delays reach 4e-6 s: keep delay * bandwidth < N, else the paths are clipped to zero power
    d = make_paths(n, 7, n_sc=max(k, 64), bandwidth=bw, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.ofdm.subcarriers = max(k, 64); p.ofdm.selected_subcarriers = np.arange(k); p.ofdm.bandwidth = bw
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    print(f"bs{bs} ue{ue} K={k} n={n} (M={m}, {8 * m * k // 1024} KB/user): " + timed(plan, out, ("mma/16:1", "mma/16:2", "mma/32:1", "mma/32:2", "tc", "ffma", "auto")), flush=True)
