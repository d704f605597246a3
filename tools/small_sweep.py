"""Small-array shapes (M <= 16) and the BASELINE cfg1 workload: fd_small2_kernel (default) against the round-1 kernels.
    python tools/small_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths, scenario
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(plan, out, variants=("auto", "small1", "ffma")):
    row = []
    for var in variants:
        if var == "auto": os.environ.pop("DMK_FD_KERNEL", None)
        else: os.environ["DMK_FD_KERNEL"] = var
        for _ in range(3): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
        row.append(f"{var}:{_lib.last_kernel().split('<')[0][3:]} {ms:.3f} ms {out.numel() * 8e-9 / (ms * 1e-3):.0f} GB/s")
    os.environ.pop("DMK_FD_KERNEL", None)
    return " | ".join(row)


for dense in (False, True):
    s = scenario(1, 80000, dense=dense)
    plan, _ = dmb.make_plan(dmb.Dataset(dict(s.data)), dmb.ChannelGenParameters(s.params), warn=False)
    out = plan.alloc_out()
    print(f"cfg1{' dense' if dense else ''} (8x1, K=64, 80k users): " + timed(plan, out), flush=True)
    print("   ", _lib.last_kernel(), flush=True)
shapes = [((8, 2), (1, 1), 64), ((4, 4), (1, 1), 64), ((16, 1), (1, 1), 1024), ((4, 2), (2, 1), 256), ((2, 2), (1, 1), 512), ((8, 1), (2, 1), 2048)]
for bs, ue, k in shapes:
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(200000, (4 << 30) // (8 * m * k)))
    bw = min(50e6, max(k, 64) / 4.2e-6)      # delays reach 4e-6 s: keep delay * bandwidth < N, else the paths are clipped to zero power
    d = make_paths(n, 7, n_sc=max(k, 64), bandwidth=bw, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.ofdm.subcarriers = max(k, 64); p.ofdm.selected_subcarriers = np.arange(k); p.ofdm.bandwidth = bw
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    print(f"bs{bs} ue{ue} K={k} n={n} (M={m}): " + timed(plan, out), flush=True)
