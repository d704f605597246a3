import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import deepmimo_b200 as dmb
from deepmimo_b200 import _lib
from deepmimo_b200.synth import make_paths
from oracle import channel_oracle as orc
from util import per_user_rel_fro
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for bs, ue, k in [((64, 1), (1, 1), 512), ((64, 4), (1, 1), 1024), ((64, 2), (2, 1), 256)]:
    m = bs[0] * bs[1] * ue[0] * ue[1]
    n = int(min(100000, (4 << 30) // (8 * m * k)))
    d = make_paths(n, 7, n_sc=1024, bandwidth=50e6, n_cols=25)
    p = dmb.ChannelGenParameters()
    p.bs_antenna.shape = np.array(bs); p.ue_antenna.shape = np.array(ue); p.bs_antenna.rotation = np.array([5, 10, 15])
    p.ofdm.subcarriers = 1024; p.ofdm.selected_subcarriers = np.arange(k); p.ofdm.bandwidth = 50e6
    plan, _ = dmb.make_plan(dmb.Dataset(d), p, warn=False)
    out = plan.alloc_out()
    res = {}
    for var in ("tc", "tc1", "ffma"):
        os.environ["DMK_FD_KERNEL"] = var
        for _ in range(2): plan.run(out)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
        for a, b in ev:
            flush.fill_(1); a.record(); plan.run(out); b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / 4
        res[var] = (ms, _lib.last_kernel().split(" ")[0], out[:64].cpu().numpy())
    o = orc.compute_channels({kk: v[:64] for kk, v in d.items()}, bs_shape=bs, ue_shape=ue, bs_rotation=[5, 10, 15], num_paths=25, subcarriers=1024,
                             selected_subcarriers=np.arange(k), bandwidth=50e6)["H"]
    print(f"bs{bs} ue{ue} K={k} n={n}: " + " | ".join(f"{v}: {res[v][1]} {res[v][0]:.3f} ms {8e-9*n*m*k/(res[v][0]*1e-3):.0f} GB/s err {per_user_rel_fro(res[v][2], o).max():.1e}" for v in res))
