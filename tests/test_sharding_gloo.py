"""N > 1 path on CPU: two gloo ranks shard two base stations' users, each computes its share with an
injected CPU compute function (the oracle -- tests may use it), and the gathered result equals the
unsharded computation.  Covers shard_plan, dataset/param slicing (per-user and random UE rotation),
FoV propagation and the summary all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import deepmimo_b200 as dmb
from cases import case_data, case_list, oracle_kwargs, params_dict
from deepmimo_b200.sharding import compute_channels_sharded, gather_summaries
from oracle import channel_oracle as orc


def _oracle_compute(ds, params, **kw):
    p = params
    return orc.compute_channels({k: ds[k] for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el")},
                                bs_shape=p.bs_antenna.shape, ue_shape=p.ue_antenna.shape, bs_spacing=p.bs_antenna.spacing,
                                ue_spacing=p.ue_antenna.spacing, bs_rotation=p.bs_antenna.rotation,
                                ue_rotation=p.ue_antenna.rotation, bs_pattern=p.bs_antenna.radiation_pattern,
                                ue_pattern=p.ue_antenna.radiation_pattern, bs_fov=ds.get("bs_fov"), ue_fov=ds.get("ue_fov"),
                                num_paths=p.num_paths, freq_domain=bool(p.freq_domain), subcarriers=p.ofdm.subcarriers,
                                selected_subcarriers=p.ofdm.selected_subcarriers, bandwidth=p.ofdm.bandwidth)["H"]


def _datasets(case_name):
    case = next(c for c in case_list() if c["name"] == case_name)
    out = []
    for b in range(2):
        c = dict(case, seed=case["seed"] + 50 * b)
        ds = dmb.Dataset(case_data(c))
        if case["bs_fov"] is not None or case["ue_fov"] is not None:
            ds.apply_fov(**{k: case[k] for k in ("bs_fov", "ue_fov") if case[k] is not None})
        out.append(ds)
    return case, out


def _worker(rank, world, port, case_name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, dss = _datasets(case_name)
        res = compute_channels_sharded(dss, dmb.ChannelGenParameters(params_dict(case)), compute=_oracle_compute)
        local = [(it.bs, it.start, it.stop, H) for it, H in res]
        gathered = gather_summaries(local)
        dist.barrier()
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case_name", ["ue_rot_random", "np10_fd_perusr_bsrot"])
def test_two_rank_gloo_sharding_equals_unsharded(case_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case, dss = _datasets(case_name)
    full = [_oracle_compute(ds, dmb.ChannelGenParameters(params_dict(case)).validate(case["n"])) for ds in dss]
    seen = [np.zeros(case["n"], int) for _ in dss]
    assert len(gathered) == 2 and all(len(g) >= 1 for g in gathered)
    for per_rank in gathered:
        for bs, a, b, H in per_rank:
            seen[bs][a:b] += 1
            assert np.array_equal(H, full[bs][a:b]), (bs, a, b)
    assert all((s == 1).all() for s in seen)
