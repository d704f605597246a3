#!/usr/bin/env python
"""bench.py -- H coefficients/s of the channel-generation hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the fused channel kernel over every user of the rank's synthetic scenario,
inputs already resident in HBM (`value`); `e2e` is the same metric through the public drop-in call
`compute_channels(dataset, params)` with HOST buffers (pinned H2D of the path matrices, kernels, D2H of
H into pinned host memory) inside the timed region.  Users are independent: ranks process disjoint
shards with no data-path collective (`scaling: weak`); NCCL is used only for the barrier and the
max-over-ranks of the device time.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference's NumPy path (oracle/channel_oracle.py,
bit-identical to the live reference on the golden cases; the reference is pure Python and cannot travel
to the GPU box) on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")

METRIC = "H coefficients/s (complex64)"
UNIT = "coef/s"
# workload name -> (BASELINE.json config number, scenario variant).  cfg1..cfg5 are SURVEY.md 8d's configurations; the others are
# sensitivity variants: every user with all 25 paths, config 3 without its FoV filter.
WORKLOADS = {"cfg1": (1, {}), "cfg2": (2, {}), "cfg3": (3, {}), "cfg4": (4, {}), "cfg5": (5, {}),
             "cfg2_dense": (2, {"dense": True}), "cfg5_dense": (5, {"dense": True}), "cfg3_nofov": (3, {"fov": False}),
             "mid_8x8_K64": (6, {}),      # 64 antennas x 64 subcarriers, 32 KB per user
             "default_8x8_K1": (7, {}),   # the reference's default OFDM parameters: one selected subcarrier
             "td_8x8_static": (8, {})}    # the reference's own time-domain mode (no time axis)
# --impl reference: users per host core in one step (about 3 s of NumPy work per core at the default K + W)
CPU_SAMPLE_USERS = {1: 12000, 2: 24, 3: 48, 4: 3000, 5: 160, 6: 1500, 7: 12000, 8: 6000}
# cpu_baseline leg of the b200 arm: one core, about 10-30 s of NumPy work
CPU_BASELINE_USERS = {1: 80000, 2: 128, 3: 256, 4: 20000, 5: 1024, 6: 10000, 7: 80000, 8: 40000}
DEFAULT_USERS = {1: 80_000, 2: 4096, 3: 8192, 4: 50_000, 5: 200_000, 6: 131_072, 7: 200_000, 8: 200_000}      # SURVEY.md 8 size table (per GPU)
FP32_LANES_PER_SM = 128
PARITY_USERS = 64              # users pulled out of the timed buffers and compared with the oracle (untimed)


def _lib_last_kernel():
    from deepmimo_b200 import _lib
    return _lib.last_kernel()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# workload bookkeeping
# ------------------------------------------------------------------------------------------------
def scenario_for(workload: str, rank: int, users):
    from deepmimo_b200.synth import scenario
    cfg, var = WORKLOADS[workload]
    if cfg == 5:
        return scenario(5, users, bs_index=rank, **var)     # one BS per GPU
    return scenario(cfg, users, shard=rank, **var)


def oracle_call(s, lo, hi):
    from oracle import channel_oracle as orc
    from util import oracle_kwargs_from_params
    data = {k: (v[lo:hi] if v.shape[0] == s.n_ue else v) for k, v in s.data.items()}
    kw = oracle_kwargs_from_params(s.params, s.bs_fov, s.ue_fov)
    rot = np.asarray(kw["ue_rotation"])
    if rot.ndim == 2 and rot.shape[0] == s.n_ue:
        kw["ue_rotation"] = rot[lo:hi]
    dop = None if s.doppler_hz is None else s.doppler_hz[lo:hi]
    return orc.compute_channels(data, **kw, doppler_hz=dop, times=s.times)["H"]


def _oracle_worker(args):
    s, lo, hi = args
    return int(oracle_call(s, lo, hi).size)


def algorithmic_counts(s, plan, info):
    """bytes = 8 n_coef + 28 n P (+24 n per-user rotation); FD flops = 8 per (coef, valid in-FoV path),
    TD flops = 6 per written slot."""
    n, p0 = plan.n_users, plan.n_cols
    shape = plan.out_shape()
    n_coef = int(np.prod(shape))
    per_user_cols = int(np.prod(shape[1:]))
    bytes_alg = 8 * n_coef + 28 * n * p0 + (24 * n if plan.ue_rot is not None else 0) + \
        (4 * n * p0 if plan.doppler is not None else 0)
    active = info.valid.copy()
    if info.fov_mask is not None:
        active &= info.fov_mask[:, :active.shape[1]]
    if plan.spec.freq_domain:
        flops = 8 * per_user_cols * int(active.sum())
    else:
        t = 1 if plan.spec.times is None else len(plan.spec.times)
        flops = 6 * plan.spec.m_rx * plan.spec.m_tx * t * int(active.sum())
    return n_coef, bytes_alg, flops, float(active.sum()) / max(n, 1)


def ring_segments(n_users: int, chunk: int, n_ring: int):
    """What the ring buffers hold after a pass over [0, n_users) in chunks of `chunk` users through `n_ring` buffers:
    a list of (first user, buffer index, first row, rows).  A short last chunk leaves the tail of the chunk that used its buffer
    before it in place -- that data was written inside the timed region too."""
    n_chunks = (n_users + chunk - 1) // chunk
    segs = []
    for b in range(min(n_ring, n_chunks)):
        i = ((n_chunks - 1 - b) // n_ring) * n_ring + b            # last chunk that went to buffer b
        rows = min(chunk, n_users - i * chunk)
        segs.append((i * chunk, b, 0, rows))
        if rows < chunk and i - n_ring >= 0:
            segs.append(((i - n_ring) * chunk + rows, b, rows, chunk - rows))
    return segs


def parity_check(s, plan, ring, chunk, info, n_sample=PARITY_USERS, seed=0):
    """Untimed: pull users out of the buffers the timed loop has just written and compare them with the oracle (the checker) --
    per-user relative Frobenius error of the coefficients against north_star's 1e-5, masks bit for bit.  The sample always holds
    the first and the last user of every ring segment, i.e. it spans the chunk boundaries of the ring."""
    import torch
    from util import TOL_REL_FRO, oracle_on_users, per_user_rel_fro
    t0 = time.perf_counter()
    segs = ring_segments(plan.n_users, chunk, len(ring))
    rng = np.random.default_rng(seed)
    where = {}
    for u0, b, r0, rows in segs:
        for u in (u0, u0 + rows - 1):
            where[u] = (b, r0 + (u - u0))
    total = sum(r for _, _, _, r in segs)
    while len(where) < min(n_sample, total):
        u0, b, r0, rows = segs[int(rng.integers(len(segs)))]
        k = int(rng.integers(rows))
        where[u0 + k] = (b, r0 + k)
    users = np.array(sorted(where))
    got = torch.stack([ring[where[u][0]][where[u][1]] for u in users]).cpu().numpy()
    o = oracle_on_users(s, users)
    err = per_user_rel_fro(got, o["H"])
    p = o["valid"].shape[1]
    masks_equal = bool(np.array_equal(info.valid[users][:, :p], o["valid"]))
    if plan.spec.freq_domain:
        masks_equal &= bool(np.array_equal(info.clip[users][:, :p], o["clip"]))
    if o["fov_mask"] is not None:
        masks_equal &= info.fov_mask is not None and bool(np.array_equal(info.fov_mask[users], o["fov_mask"]))
    nan_free = not bool(np.isnan(got.view(np.float32)).any())
    return {"max_rel_fro": float(err.max()), "tolerance": TOL_REL_FRO, "masks_equal": masks_equal, "users": int(len(users)),
            "ring_segments": len(segs), "nan_free": nan_free,
            "ok": bool(err.max() <= TOL_REL_FRO and masks_equal and nan_free),
            "checker": "oracle/channel_oracle.py on users pulled from the timed output buffers", "seconds": round(time.perf_counter() - t0, 2)}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int, interval=0.02):
        super().__init__(daemon=True)
        self.interval, self.samples, self.reasons, self.power = interval, [], set(), []
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # prime every query once here, outside the timed region: the first call of an NVML query can take milliseconds
            # and was seen to stretch one timed step from 2.9 to 6.2 ms
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetPowerUsage(self.h)
            (pynvml.nvmlDeviceGetCurrentClocksEventReasons if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons")
             else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(self.h)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            log(f"[bench] NVML unavailable ({e}); clocks not sampled")

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.interval)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_min_mhz": float(min(self.samples)),
                "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
def time_plan(plan, out_ring, chunk, steps, warmup, flush, dist, world):
    """Per-step CUDA-event timing on the launch stream; L2 flushed (untimed) between steps.
    Returns (sum of step times in ms on this rank, list of step ms)."""
    import torch
    from deepmimo_b200.channels import chunk_is_independent
    stream = torch.cuda.current_stream()
    n = plan.n_users

    def one_step():
        i = 0
        for a in range(0, n, chunk):
            b = min(a + chunk, n)
            plan.run(out_ring[i % len(out_ring)][: b - a], a, b,
                     independent=chunk_is_independent(i, len(out_ring)) and not os.environ.get("DMK_BENCH_NO_PDL"))
            i += 1

    for _ in range(warmup):
        one_step()
        flush.fill_(1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k)                      # > L2 (126 MB): evict inputs/outputs of the previous step
        ev[k][0].record(stream)
        one_step()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    return float(sum(ms)), ms


def run_workload(workload, users, steps, warmup, rank, world, dist, flush, want_clocks=True, parity=True):
    """One workload on this rank.  The rank's share comes from the product's sharding path
    (deepmimo_b200.sharding.compute_channels_sharded over one dataset per base station / shard, lazily materialised), with
    the timed ring loop as the `compute` consumer."""
    import torch
    import deepmimo_b200 as dmb
    from deepmimo_b200 import _lib
    from deepmimo_b200.channels import default_chunk_users
    from deepmimo_b200.sharding import compute_channels_sharded

    cfg = WORKLOADS[workload][0]
    n_per = users if users is not None else DEFAULT_USERS[cfg]
    made = {rank: scenario_for(workload, rank, users)}          # this rank's base station / shard; the others are never built here

    def factory(b):
        def make():
            s = made[b] if b in made else made.setdefault(b, scenario_for(workload, b, users))
            ds = dmb.Dataset(dict(s.data))
            if s.bs_fov is not None:
                ds.apply_fov(bs_fov=s.bs_fov, ue_fov=s.ue_fov)
            return ds
        return make

    out = {}

    def timed(ds, params):
        s = made[rank]
        plan, _ = dmb.make_plan(ds, params, times=s.times, doppler=s.doppler_hz, warn=False)
        per_user = plan.spec.coefs_per_user(plan.n_cols) * 8
        total_bytes = per_user * plan.n_users
        free_b, _tot = torch.cuda.mem_get_info()
        if total_bytes <= min(64 << 30, int(free_b * 0.6)) and not os.environ.get("DMK_BENCH_FORCE_RING"):
            chunk, ring = plan.n_users, [plan.alloc_out()]
            layout = f"single [{plan.n_users} users] output tensor ({total_bytes / 2**30:.1f} GiB) rewritten every step"
        else:
            chunk = default_chunk_users(plan, int(float(os.environ.get("DMK_BENCH_CHUNK_GIB", "4")) * (1 << 30)))   # ring of 4 GiB chunks (SURVEY.md 8d cfg 5)
            n_ring = int(os.environ.get("DMK_BENCH_RING", "3"))
            ring = [plan.alloc_out(chunk) for _ in range(n_ring)]
            layout = f"ring of {n_ring} x {chunk} users ({chunk * per_user / 2**30:.1f} GiB) output chunks in HBM"
        # masks once (also gives the algorithmic flop count)
        masks = plan.alloc_masks()
        i = 0
        for a in range(0, plan.n_users, chunk):
            b = min(a + chunk, plan.n_users)
            plan.run(ring[i % len(ring)][: b - a], a, b, {k: v[a:b] for k, v in masks.items()})
            i += 1
        torch.cuda.synchronize()
        info = plan.info_from_masks(masks)
        n_coef, bytes_alg, flops, pbar = algorithmic_counts(s, plan, info)
        for r in ring:
            r.fill_(float("nan"))                               # whatever parity_check reads below was written by the timed loop

        sampler = ClockSampler(torch.cuda.current_device()) if (want_clocks and not os.environ.get("DMK_BENCH_NO_CLOCKS")) else None
        l0 = _lib.launch_count()
        if sampler:
            sampler.start()
        total_ms, ms = time_plan(plan, ring, chunk, steps, warmup, flush, dist, world)
        if sampler:
            sampler.stop_flag.set()
            sampler.join()
        launches = (_lib.launch_count() - l0) * steps // (steps + warmup)      # launches inside the timed region
        par = None
        if parity:
            try:
                par = parity_check(s, plan, ring, chunk, info, seed=rank)
            except Exception as e:  # noqa: BLE001
                par = {"ok": False, "error": str(e)[:300]}
        out.update(scenario=s, plan=plan, ds=ds, params=params, total_ms=total_ms, ms=ms, n_coef=n_coef, bytes_alg=bytes_alg,
                   flops=flops, pbar=pbar, layout=layout, launches=launches, kernel=_lib.last_kernel(), parity=par,
                   clocks=sampler.result() if sampler else None, launches_per_step=(plan.n_users + chunk - 1) // chunk)
        return None

    items = compute_channels_sharded([factory(b) for b in range(world)], dmb.ChannelGenParameters(made[rank].params), rank=rank,
                                     world_size=world, compute=timed, sizes=[n_per] * world)
    assert len(items) == 1 and items[0][0].bs == rank and items[0][0].n == n_per, items
    out["shard"] = f"bs/shard {items[0][0].bs}, users [{items[0][0].start}, {items[0][0].stop})"
    return out


def run_e2e(res, steps, rank, world, dist, cap_bytes=4 << 30):
    """Public-API path with host buffers, every step: pinned H2D of the path matrices + kernels + D2H of H.  Three legs on the
    same users: `pinned` (caller-provided page-locked result buffer), `default` (the call a user makes, no host_out: the result
    is allocated inside), and `ceiling` (the D2H copies alone into the same pinned buffer in the same chunks, no kernels) --
    what the host link of this box delivers when every rank copies at once."""
    import torch
    import deepmimo_b200 as dmb
    from deepmimo_b200.channels import default_chunk_users
    s, plan = res["scenario"], res["plan"]
    per_user = plan.spec.coefs_per_user(plan.n_cols) * 8
    n = int(max(1, min(plan.n_users, cap_bytes // per_user)))

    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    data = {k: (pinned(v[:n]) if v.shape[0] == s.n_ue else v) for k, v in s.data.items()}
    prm = dmb.ChannelGenParameters(s.params)
    rot = np.asarray(prm.ue_antenna.rotation)
    if rot.ndim == 2 and rot.shape[0] == s.n_ue:
        prm.ue_antenna.rotation = pinned(rot[:n])
    dop = None if s.doppler_hz is None else pinned(s.doppler_hz[:n])
    ds = dmb.Dataset(data)
    if s.bs_fov is not None:
        ds.apply_fov(bs_fov=s.bs_fov, ue_fov=s.ue_fov)
    shape = plan.spec.out_shape(n, plan.n_cols)
    host_out = torch.empty(shape, dtype=torch.complex64, pin_memory=True)
    h2d = sum(int(data[k].nbytes) for k in ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el"))
    h2d += (int(np.asarray(prm.ue_antenna.rotation).nbytes) if rot.ndim == 2 else 0) + (int(dop.nbytes) if dop is not None else 0)
    d2h = int(np.prod(shape)) * 8

    def call_pinned():
        return dmb.compute_channels(ds, prm, times=s.times, doppler=dop, host_out=host_out, cache=False, warn=False)

    def call_default():
        return dmb.compute_channels(ds, prm, times=s.times, doppler=dop, warn=False)     # result allocated inside, cached on ds

    sub_plan, _ = dmb.make_plan(ds, prm, times=s.times, doppler=dop, warn=False)
    chunk = default_chunk_users(sub_plan)
    dev_bufs = [sub_plan.alloc_out(min(chunk, n)) for _ in range(2 if n > chunk else 1)]
    for b in dev_bufs:
        b.zero_()
    copy = torch.cuda.Stream()

    def call_ceiling():
        with torch.cuda.stream(copy):
            for i, a in enumerate(range(0, n, chunk)):
                z = min(a + chunk, n)
                host_out[a:z].copy_(dev_bufs[i % len(dev_bufs)][: z - a], non_blocking=True)
        copy.synchronize()

    def timed(fn, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()                            # returns after the copy stream has drained (host array complete)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        return dt

    t_pin = timed(call_pinned, 2)
    t_ceil = timed(call_ceiling, 1)
    del host_out                            # back to torch's pinned-block cache: the default leg's first result reuses it
    t_def = timed(call_default, 3)          # the first two calls page-lock their result blocks; afterwards they are reused
    ds._data.pop("channel", None)
    return dict(seconds=t_pin, seconds_default=t_def, seconds_ceiling=t_ceil, steps=steps, users=n,
                coefs_per_step=int(np.prod(shape)), h2d=h2d, d2h=d2h)


def cpu_baseline(s, workload, cores=1):
    cfg = WORKLOADS[workload][0]
    n = min(s.n_ue, CPU_BASELINE_USERS[cfg])
    t0 = time.perf_counter()
    H = oracle_call(s, 0, n)
    dt = time.perf_counter() - t0
    return {"value": H.size / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} users of {workload} ({H.size} coefficients) in {dt:.1f} s, oracle/channel_oracle.py "
                      f"(NumPy {np.__version__}, single thread like the reference's per-user loop)"}


def bind_to_gpu_cpus(device_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root) before any pinned host
    buffer is allocated: with one process per GPU the D2H copies of the e2e leg then land in node-local memory instead
    of crossing the socket interconnect.  Returns the CPU count used, or None if NVML / the syscall is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as e:  # noqa: BLE001
        log(f"[bench] CPU affinity not set ({e})")
    return None


def gpu_main(args):
    # stdout must carry exactly one JSON line: libraries (NCCL prints its version banner) write to fd 1, so
    # fd 1 is pointed at stderr for the run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the b200 arm")
    torch.cuda.set_device(local)
    n_local_cpus = bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    users = args.users

    res = run_workload(args.workload, users, args.steps, args.warmup, rank, world, dist, flush)
    t_local = torch.tensor([res["total_ms"]], dtype=torch.float64, device="cuda")
    coefs = torch.tensor([float(res["n_coef"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        dist.all_reduce(coefs, op=dist.ReduceOp.SUM)
    total_ms, total_coefs = float(t_local.item()), float(coefs.item())
    value = total_coefs * args.steps / (total_ms / 1e3)

    e2e = run_e2e(res, args.e2e_steps, rank, world, dist)

    def agg_rate(seconds, coefs_per_step, nsteps):
        e_t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
        e_c = torch.tensor([float(coefs_per_step)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(e_c, op=dist.ReduceOp.SUM)
        return float(e_c.item()) * nsteps / float(e_t.item())

    e2e_value = agg_rate(e2e["seconds"], e2e["coefs_per_step"], e2e["steps"])
    e2e_default = agg_rate(e2e["seconds_default"], e2e["coefs_per_step"], e2e["steps"])
    e2e_ceiling = agg_rate(e2e["seconds_ceiling"], e2e["coefs_per_step"], e2e["steps"])

    def summarise(r, nsteps):
        """Per-workload record; times are the max over ranks, rates the whole-job aggregate."""
        t = torch.tensor([r["total_ms"]], dtype=torch.float64, device="cuda")
        c = torch.tensor([float(r["n_coef"]), float(r["bytes_alg"]), float(r["flops"])], dtype=torch.float64, device="cuda")
        ok = torch.tensor([1.0 if (r["parity"] or {}).get("ok") else 0.0, (r["parity"] or {}).get("max_rel_fro", float("nan"))],
                          dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            okmin, errmax = ok[:1].clone(), ok[1:].clone()
            dist.all_reduce(okmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(errmax, op=dist.ReduceOp.MAX)
            ok = torch.cat([okmin, errmax])
        ms = float(t.item()) / nsteps
        n_coef, bytes_alg, flops = (float(v) for v in c.tolist())
        gbs_per_gpu = bytes_alg / world / (ms / 1e3) / 1e9
        fp32_peak = 2 * 148 * FP32_LANES_PER_SM * sm_max_mhz * 1e6 / 1e12
        t_write, t_fma = bytes_alg / world / (hbm_peak * 1e9), flops / world / (fp32_peak * 1e12)
        par = dict(r["parity"] or {})
        if world > 1:
            par.update(ok=bool(ok[0].item() > 0.5), max_rel_fro=float(ok[1].item()), ranks=world)
        return {"coef_per_s": n_coef / (ms / 1e3), "ms_per_step": ms, "users_per_gpu": r["plan"].n_users,
                "gb_per_s_per_gpu": gbs_per_gpu, "hbm_frac": gbs_per_gpu / hbm_peak,
                "tflops_fp32_per_gpu": flops / world / (ms / 1e3) / 1e12, "t_min_over_t": max(t_write, t_fma) / (ms / 1e3),
                "t_min_bound": "fp32" if t_fma > t_write else "hbm", "mean_active_paths": r["pbar"], "layout": r["layout"],
                "kernel": r["kernel"].split(" ")[0], "shard": r["shard"], "parity": par}

    extra = {}
    if args.others:
        # the city-scale configuration runs at EVERY N (one base station per rank through the sharding path); the remaining
        # shapes and the sensitivity variants only at N = 1
        names = ["cfg5"] if world > 1 else ["cfg1", "cfg3", "cfg4", "cfg5", "cfg2_dense", "cfg5_dense", "cfg3_nofov", "mid_8x8_K64", "default_8x8_K1", "td_8x8_static"]
        for w in [w for w in names if w != args.workload]:
            try:
                r = run_workload(w, None, 5, 3, rank, world, dist, flush, want_clocks=True)
                rec = summarise(r, 5)
                rec["clocks"] = r["clocks"]
                if rank == 0:
                    extra[w] = rec
                del r
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                if world > 1:
                    raise
                extra[w] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and args.others:
        # row f3: fused beam amplitude map (16-beam steering_vec codebook) on the headline shape -- H is never written
        try:
            import deepmimo_b200 as dmb
            plan = res["plan"]
            if plan.spec.freq_domain and plan.spec.times is None:
                F = np.array([dmb.steering_vec(list(plan.spec.bs_shape), phi=a, spacing=plan.spec.bs_spacing).squeeze()
                              for a in np.around(np.linspace(-60, 60, 16), 2)]).astype(np.complex64)
                Fd = torch.from_numpy(F).cuda()
                amp = torch.empty((plan.n_users, 16), dtype=torch.float32, device="cuda")
                for _ in range(3):
                    plan.run_beams(Fd, amp)
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
                for a, b in evs:
                    flush.fill_(3)
                    a.record(); plan.run_beams(Fd, amp); b.record()
                torch.cuda.synchronize()
                ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
                extra["beams16_" + args.workload] = {
                    "ms_per_step": ms, "users": plan.n_users, "users_per_s": plan.n_users / (ms / 1e3),
                    "beamformed_coef_per_s": plan.n_users * plan.spec.m_rx * 16 * len(plan.spec.selected) / (ms / 1e3),
                    "equivalent_H_coef_per_s": res["n_coef"] / (ms / 1e3), "kernel": _lib_last_kernel().split(" ")[0],
                    "what": "mean_{r,k} |F @ H| for a 16-beam codebook straight from the path matrices (dmk_beam_amplitude_fd)"}
        except Exception as e:  # noqa: BLE001
            extra["beams16_" + args.workload] = {"error": str(e)[:200]}

    if rank == 0:
        ms_step = total_ms / args.steps
        n_launch = res["launches_per_step"]
        t_kernel = (res["total_ms"] / args.steps) / 1e3 / n_launch        # average launch duration on rank 0
        gbps = res["bytes_alg"] / n_launch / t_kernel / 1e9
        clocks = res["clocks"] or {}
        f_clk = sm_max_mhz * 1e6
        fp32_peak = 2 * 148 * FP32_LANES_PER_SM * f_clk / 1e12
        tfl = res["flops"] / n_launch / t_kernel / 1e12
        t_write = res["bytes_alg"] / (hbm_peak * 1e9)
        t_fma = res["flops"] / (fp32_peak * 1e12)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                rec = json.load(f).get(args.workload)
            if rec:      # measured at rec["users"] users per launch; DRAM traffic is linear in users (every user is written once)
                traffic = rec["dram_bytes"] * (res["plan"].n_users / rec["users"]) / n_launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_step_min_median_max": [min(res["ms"]), statistics.median(res["ms"]), max(res["ms"])],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {res['scenario'].notes}", "users_per_gpu": res["plan"].n_users,
                       "out_shape_per_gpu": list(res["plan"].out_shape()), "mean_active_paths_per_user": res["pbar"],
                       "sharding": "independent (BS, user) shards per GPU, no data-path collective",
                       "output": res["layout"],
                       "l2": "512 MiB flush write between timed steps (untimed); output per step >> 126 MB L2",
                       "timing": "per-step CUDA events on the launch stream, summed; max over ranks"},
            "roofline": {"bound": "hbm", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s", "frac": gbps / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": res["kernel"],
                         "algorithmic_bytes_per_launch": res["bytes_alg"] / n_launch,
                         "fp32": {"achieved_tflops": tfl, "peak_tflops": fp32_peak, "frac": tfl / fp32_peak,
                                  "peak_source": f"2*148 SM*128 lanes*{sm_max_mhz:.0f} MHz (nominal max clock)",
                                  "algorithmic_flops_per_launch": res["flops"] / n_launch},
                         "t_min_over_t": max(t_write, t_fma) / (ms_step / 1e3 * (1 if world == 1 else 1)),
                         "t_min_bound": "fp32" if t_fma > t_write else "hbm"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "users_per_gpu": e2e["users"], "steps": e2e["steps"],
                    "ceiling": e2e_ceiling, "frac": e2e_value / e2e_ceiling,
                    "ceiling_what": "the same D2H copies (same pinned buffer, same chunks, all ranks at once) with no kernels: "
                                    "what this box's host link delivers",
                    "d2h_gb_per_s": e2e_value * 8 / 1e9, "ceiling_gb_per_s": e2e_ceiling * 8 / 1e9,
                    "cpu_affinity": (f"rank pinned to the {n_local_cpus} CPUs NVML reports local to its GPU" if n_local_cpus else "not set"),
                    "path": "deepmimo_b200.compute_channels(dataset, params, host_out=pinned): pinned H2D + fused kernel "
                            "(1 GiB chunks, 2 device buffers) + D2H overlapped on a copy stream"},
            "e2e_default": {"value": e2e_default, "unit": UNIT, "frac_of_pinned": e2e_default / e2e_value,
                            "path": "deepmimo_b200.compute_channels(dataset, params) -- no host_out: the result is allocated inside "
                                    "(page-locked up to DMK_PINNED_CAP_GIB, blocks of dropped results are reused) and cached on the dataset"},
            "parity": res["parity"],
            "gpu_launches": res["launches"],
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(res["scenario"], args.workload)
        if extra:
            line["workloads"] = extra
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# the reference (CPU) arm
# ------------------------------------------------------------------------------------------------
def reference_main(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = WORKLOADS[args.workload][0]
    cores = os.cpu_count() or 1
    # bounded sample: scale the per-step sample so that K + W steps end within a few minutes
    per = max(1, int(CPU_SAMPLE_USERS[cfg] * min(1.0, 15.0 / (args.steps + args.warmup))))
    s = scenario_for(args.workload, 0, per * cores)
    jobs = [(s, i * per, (i + 1) * per) for i in range(cores)]
    ctx = mp.get_context("fork")
    times, coefs = [], 0
    with ctx.Pool(cores) as pool:
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            sizes = pool.map(_oracle_worker, jobs)
            dt = time.perf_counter() - t0
            if k >= args.warmup:
                times.append(dt)
                coefs = sum(sizes)
    total = sum(times)
    value = coefs * len(times) / total
    sample = (f"{per * cores} users of {args.workload} per step ({coefs} coefficients), {cores} processes x {per} users, "
              f"oracle/channel_oracle.py (NumPy {np.__version__}); the reference itself is single-threaded")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/c128 (NumPy)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {s.notes}", "sample_users_per_step": per * cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--users", type=int, default=None, help="users per GPU (default: the configuration's size)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--others", action="store_true", default=True, help="also time the other configurations briefly (N=1)")
    ap.add_argument("--no-others", dest="others", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    return gpu_main(args)


if __name__ == "__main__":
    main()
