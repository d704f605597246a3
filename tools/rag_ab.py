import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for bs, ue, k in (((32, 8), (2, 2), 100), ((32, 8), (2, 2), 88), ((32, 8), (1, 1), 150), ((16, 8), (1, 1), 340), ((8, 8), (1, 1), 270)):
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(200000, (4 << 30) // (8 * m * k)))
    n_sc = int(2 ** np.ceil(np.log2(k))); bw = min(50e6, n_sc / 4.2e-6)
    d = make_paths(n, 7, n_sc=n_sc, bandwidth=bw, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.ofdm.bandwidth = bw
    p.ofdm.subcarriers = n_sc; p.ofdm.selected_subcarriers = np.arange(k)
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out(); row = []
    for var in ("auto", "tc", "ffma", "mma"):
        if var == "auto": os.environ.pop("DMK_FD_KERNEL", None)
        else: os.environ["DMK_FD_KERNEL"] = var
        for _ in range(3): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
        row.append(f"{var}:{_lib.last_kernel().split('<')[0][3:]} {out.numel() * 8e-9 / (ms * 1e-3):.0f} GB/s")
    os.environ.pop("DMK_FD_KERNEL", None)
    print(f"bs{bs} ue{ue} K={k} (M={m}, {8*m*k//1024} KB/user): " + " | ".join(row), flush=True)
