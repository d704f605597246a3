"""Multi-GPU sharding of the channel path: independent (BS, user) units, no data-path collective.

Every (base station, user) pair is independent in the reference (deepmimo/generator/channel.py:264-287
touches only row i; base stations are separate Datasets, dataset.py:947-950), so the work is a flat list
of users across base stations that is cut into `world_size` contiguous ranges.  One process per GPU
(torchrun); torch.distributed is used only for rendezvous, barriers and for gathering small per-shard
summaries -- H stays on the GPU that produced it.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class ShardItem:
    bs: int        # index of the base-station dataset
    start: int     # first user (inclusive)
    stop: int      # last user (exclusive)

    @property
    def n(self) -> int:
        return self.stop - self.start


def shard_plan(sizes: Sequence[int], world_size: int) -> List[List[ShardItem]]:
    """Cut the flattened (BS, user) list into `world_size` contiguous, near-equal ranges.

    sizes[b] = number of users of base station b.  Rank r gets a list of (bs, start, stop) items; the
    union over ranks covers every user exactly once, ranks differ by at most one user.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    total = int(sum(sizes))
    bounds = [(total * r) // world_size for r in range(world_size + 1)]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    plan: List[List[ShardItem]] = []
    for r in range(world_size):
        lo, hi = bounds[r], bounds[r + 1]
        items = []
        for b, n in enumerate(sizes):
            a, z = max(lo, int(offsets[b])), min(hi, int(offsets[b + 1]))
            if z > a:
                items.append(ShardItem(b, a - int(offsets[b]), z - int(offsets[b])))
        plan.append(items)
    return plan


def rank_world(rank: Optional[int] = None, world_size: Optional[int] = None):
    if rank is None or world_size is None:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except Exception:  # noqa: BLE001
            pass
        return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world_size


def slice_dataset(ds, start: int, stop: int):
    """A Dataset view of users [start, stop): every per-user array (first dimension n_ue) is sliced."""
    from .dataset import Dataset
    n = int(np.asarray(ds["power"]).shape[0])
    out = {}
    for k in list(ds.keys()):
        v = ds._data[k] if hasattr(ds, "_data") else ds[k]
        if k in ("channel", "ch_params") or str(k).startswith("_"):
            continue
        out[k] = v[start:stop] if hasattr(v, "shape") and len(getattr(v, "shape", ())) >= 1 and v.shape[0] == n and k != "tx_pos" else v
    return Dataset(out)


def slice_params(params, n_ue: int, start: int, stop: int):
    """Per-user UE rotation follows the user slice; everything else is shared."""
    p = params.deepcopy()
    rot = p["ue_antenna"].get("rotation")
    if rot is not None:
        rot = np.asarray(rot)
        if rot.ndim == 2 and rot.shape == (n_ue, 3):
            p["ue_antenna"]["rotation"] = rot[start:stop]
    return p


def compute_channels_sharded(datasets, params, *, rank: Optional[int] = None, world_size: Optional[int] = None,
                             compute: Optional[Callable] = None, sizes: Optional[Sequence[int]] = None, **kwargs):
    """Compute this rank's share of the channels of one or several base-station datasets.

    `datasets`: a Dataset, a MacroDataset or a list of Datasets.  A list entry may also be a zero-argument callable that
    returns the Dataset: with `sizes` (users per base station) given, a rank only materialises the base stations it owns a
    part of -- at city scale (8 x 200 k users) no rank loads the other ranks' ray data.  Returns a list of
    (ShardItem, H) for the calling rank, H as returned by `compute(dataset_slice, params_slice, **kwargs)` (default:
    deepmimo_b200.compute_channels with out='torch', i.e. a CUDA tensor that stays on this GPU; a streaming consumer passes
    its own function, e.g. one that walks `iter_channels`).
    A (3,2) random UE rotation is drawn for the whole base station first (same values as the
    unsharded call) and then sliced.
    """
    from .channels import compute_channels as _cc, resolve_ue_rotation
    if hasattr(datasets, "datasets"):
        datasets = datasets.datasets
    elif not isinstance(datasets, (list, tuple)):
        datasets = [datasets]
    datasets = list(datasets)
    rank, world_size = rank_world(rank, world_size)
    if sizes is None:
        datasets = [d() if callable(d) and not hasattr(d, "keys") else d for d in datasets]
        sizes = [int(np.asarray(d["power"]).shape[0]) for d in datasets]
    elif len(sizes) != len(datasets):
        raise ValueError("sizes must have one entry per dataset")
    plan = shard_plan(sizes, world_size)[rank]
    if compute is None:
        kwargs.setdefault("out", "torch")
        kwargs.setdefault("cache", False)
        compute = _cc
    results = []
    for it in plan:
        ds = datasets[it.bs]
        if callable(ds) and not hasattr(ds, "keys"):
            ds = datasets[it.bs] = ds()
        n = sizes[it.bs]
        if int(np.asarray(ds["power"]).shape[0]) != n:
            raise ValueError(f"dataset {it.bs} has {int(np.asarray(ds['power']).shape[0])} users, sizes says {n}")
        p = params.deepcopy()
        p.validate(n)
        rot = p["ue_antenna"].get("rotation")
        if rot is not None and np.asarray(rot).shape == (3, 2):
            np.random.seed(1001)
            _, per_user = resolve_ue_rotation(rot, n, seed_numpy_rng=False)
            p["ue_antenna"]["rotation"] = per_user
        whole = it.start == 0 and it.stop == n
        sub = ds if whole else slice_dataset(ds, it.start, it.stop)
        if not whole:
            for key in ("bs_fov", "ue_fov"):
                v = ds.get(key) if hasattr(ds, "get") else None
                if v is not None:
                    sub[key] = v
        results.append((it, compute(sub, p if whole else slice_params(p, n, it.start, it.stop), **kwargs)))
    return results


def gather_summaries(local: list, group=None) -> list:
    """all_gather of small picklable per-rank summaries (counts, checksums, timings)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local, group=group)
    return out
