"""Drop-in `compute_channels` on B200: host driver over the libdmk C ABI.

Replaces the reference call `Dataset.compute_channels(params)` (deepmimo/generator/dataset.py:224-268)
for both this package's `Dataset` and an unmodified reference `dm.Dataset` (duck-typed: item access,
`.get`, optional `.set_channel_params`).  The host side keeps what the reference keeps on the host:
parameter validation (channel.py:78-139), the UE random-rotation draw from NumPy's global RNG after
`np.random.seed(1001)` (dataset.py:250,:332-338), the delay-overflow warning (channel.py:228-250),
chunk scheduling and copies.  Everything numeric runs in the CUDA kernels; there is no CPU fallback.

Layers
  parse_spec()      ChannelGenParameters (+FoV, times) -> ChannelSpec        (pure host logic, CPU-testable)
  ChannelPlan       device-resident path matrices + packed dmk_desc; .run() launches one user range
  compute_channels  the public call: H2D -> kernels (chunked, double-buffered) -> D2H / CUDA tensor
  iter_channels     chunk iterator over a ring of device buffers (streaming when H exceeds HBM)
"""
from __future__ import annotations

import ctypes
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Iterator, Optional, Tuple

import numpy as np

from . import _lib
from .params import ChannelGenParameters, RADIATION_PATTERNS

_RNG_LOCK = threading.Lock()
PATH_KEYS = ("power", "phase", "delay", "aoa_az", "aoa_el", "aod_az", "aod_el")   # deepmimo/consts.py:188-194
MAX_COLS = 32   # DMK_MAX_PATHS


# ------------------------------------------------------------------------------------------------
# host logic
# ------------------------------------------------------------------------------------------------
@dataclass
class ChannelSpec:
    """Numeric content of ChannelGenParameters + FoV + time axis for one compute_channels call."""
    bs_shape: Tuple[int, int]
    ue_shape: Tuple[int, int]
    bs_spacing: float
    ue_spacing: float
    bs_rot: np.ndarray                       # [3] degrees
    ue_rot: np.ndarray                       # [3] degrees (uniform) -- ignored when ue_rot_users is set
    ue_rot_users: Optional[np.ndarray]       # [n,3] float64 degrees or None
    bs_fov: Optional[np.ndarray]
    ue_fov: Optional[np.ndarray]
    fov_any: bool
    fov_side: Tuple[bool, bool]
    patterns: Tuple[int, int]
    num_paths: int
    freq_domain: bool
    n_subcarriers: int
    selected: np.ndarray                     # int32 [K]
    subc_start: int
    subc_step: int                           # 0 = not affine
    bandwidth: float
    rx_filter: int
    times: Optional[np.ndarray]              # float64 [T] or None
    extra: dict = field(default_factory=dict)

    @property
    def m_tx(self) -> int:
        return self.bs_shape[0] * self.bs_shape[1]

    @property
    def m_rx(self) -> int:
        return self.ue_shape[0] * self.ue_shape[1]

    def n_paths_eff(self, n_cols: int) -> int:
        return min(self.num_paths, n_cols)

    def out_shape(self, n_users: int, n_cols: int) -> Tuple[int, ...]:
        last = len(self.selected) if self.freq_domain else self.n_paths_eff(n_cols)
        shp = (n_users, self.m_rx, self.m_tx, last)
        return shp if self.times is None else shp + (len(self.times),)

    def coefs_per_user(self, n_cols: int) -> int:
        return int(np.prod(self.out_shape(1, n_cols)))


def _is_full_fov(fov) -> bool:
    """deepmimo/generator/dataset.py:450-459."""
    return bool(fov[0] >= 360 and fov[1] >= 180)


def resolve_ue_rotation(rot, n_ue: int, seed_numpy_rng: bool = True):
    """UE rotation -> (uniform [3], per_user [n,3] | None).  dataset.py:328-338.

    (3,) constant; (3,2) = per-axis [lo, hi] range drawn per user from NumPy's *global* RNG right
    after np.random.seed(1001) (dataset.py:250) so the draw is bit-identical to the reference's;
    (n,3) per-user values.
    """
    rot = np.asarray(rot)
    if rot.ndim == 1 and rot.shape[0] == 3:
        return rot.astype(np.float64), None
    if rot.ndim == 2 and rot.shape == (3, 2):
        if seed_numpy_rng:
            np.random.seed(1001)
        return np.zeros(3), np.random.uniform(rot[:, 0], rot[:, 1], (n_ue, 3))
    if rot.ndim == 2 and rot.shape == (n_ue, 3):
        # float64 like the reference's typical inputs; a float32 array would take a float32
        # dtype flow in NumPy (deg2rad/sin in float32) that this path does not reproduce.
        return np.zeros(3), np.ascontiguousarray(rot, dtype=np.float64)
    raise ValueError(f"UE rotation has shape {rot.shape}; expected (3,), (3, 2) or ({n_ue}, 3)")


def parse_spec(params, n_ue: int, *, bs_fov=None, ue_fov=None, times=None, seed_numpy_rng: bool = True) -> ChannelSpec:
    """Translate (validated) channel parameters into a ChannelSpec.  Raises like the reference does:
    NotImplementedError for an unknown pattern (ant_patterns.py:119-122).  `enable_dual_polar` is accepted and has no
    effect, exactly as in the reference (the key exists in the defaults, channel.py:51, and nothing reads it);
    `ofdm.rx_filter=1` selects the receive low-pass filter (channel.py:193-194) in the frequency-domain branch."""
    bs, ue, ofdm = params["bs_antenna"], params["ue_antenna"], params["ofdm"]
    pats = []
    for side, name in ((bs, "TX"), (ue, "RX")):
        pat = side["radiation_pattern"]
        if pat not in RADIATION_PATTERNS:
            raise NotImplementedError(f"The given '{pat}' antenna radiation pattern is not applicable for {name}.")
        pats.append(RADIATION_PATTERNS.index(pat))

    bs_shape = tuple(int(v) for v in np.asarray(bs["shape"]).ravel()[:2])
    ue_shape = tuple(int(v) for v in np.asarray(ue["shape"]).ravel()[:2])
    if len(bs_shape) != 2 or len(ue_shape) != 2:
        raise ValueError("antenna shape needs at least two entries")
    bs_rot = bs.get("rotation")
    bs_rot = np.zeros(3) if bs_rot is None else np.asarray(bs_rot, dtype=np.float64)
    if bs_rot.shape != (3,):
        raise ValueError("The BS antenna rotation must be a 3D vector")
    ue_rot, ue_rot_users = resolve_ue_rotation(ue.get("rotation") if ue.get("rotation") is not None else np.zeros(3),
                                               n_ue, seed_numpy_rng)

    # FoV: Dataset.apply_fov values win; `fov` inside the antenna dicts is the alternative spelling.
    if bs_fov is None and bs.get("fov") is not None:
        bs_fov = bs.get("fov")
    if ue_fov is None and ue.get("fov") is not None:
        ue_fov = ue.get("fov")
    bs_fov = None if bs_fov is None else np.asarray(bs_fov, dtype=np.float64)
    ue_fov = None if ue_fov is None else np.asarray(ue_fov, dtype=np.float64)
    bs_full = bs_fov is None or _is_full_fov(bs_fov)
    ue_full = ue_fov is None or _is_full_fov(ue_fov)
    fov_any = not (bs_full and ue_full)                    # dataset.py:484
    fov_side = (fov_any and not bs_full, fov_any and not ue_full)

    sel = np.asarray(ofdm["selected_subcarriers"]).ravel()
    if sel.size and not np.all(sel == np.round(sel)):
        raise ValueError("selected_subcarriers must be integers")
    sel = sel.astype(np.int32)
    start, step = 0, 0
    if sel.size == 1:
        start, step = int(sel[0]), 1
    elif sel.size > 1:
        dif = np.diff(sel.astype(np.int64))
        if np.all(dif == dif[0]) and dif[0] != 0:
            start, step = int(sel[0]), int(dif[0])
    tt = None if times is None else np.ascontiguousarray(np.atleast_1d(np.asarray(times, dtype=np.float64)))
    return ChannelSpec(bs_shape=bs_shape, ue_shape=ue_shape, bs_spacing=float(bs["spacing"]), ue_spacing=float(ue["spacing"]),
                       bs_rot=bs_rot, ue_rot=ue_rot, ue_rot_users=ue_rot_users, bs_fov=bs_fov, ue_fov=ue_fov,
                       fov_any=fov_any, fov_side=fov_side, patterns=(pats[0], pats[1]),
                       num_paths=int(params["num_paths"]), freq_domain=bool(params["freq_domain"]),
                       n_subcarriers=int(ofdm["subcarriers"]), selected=sel, subc_start=start, subc_step=step,
                       bandwidth=float(ofdm["bandwidth"]), rx_filter=int(ofdm.get("rx_filter", 0)), times=tt)


def delay_overflow_warning(delay: np.ndarray, spec: ChannelSpec, n_cols_used: int) -> bool:
    """The reference's printed warning when a delay exceeds the OFDM symbol (channel.py:228-250)."""
    if not spec.freq_domain or delay.size == 0:
        return False
    d = delay[:, :n_cols_used]
    finite = d[~np.isnan(d)]
    if finite.size == 0:
        return False
    max_delay = float(finite.max())
    n, b = spec.n_subcarriers, spec.bandwidth
    sym = n * (1 / b)
    if not max_delay > sym:
        return False
    bar = "-" * 50
    print("\nWarning: Some path delays exceed OFDM symbol duration")
    print(bar)
    print("OFDM Configuration:")
    print(f"- Number of subcarriers (N): {n}")
    print(f"- Bandwidth (B): {b / 1e6:.1f} MHz")
    print(f"- Subcarrier spacing (Δf = B/N): {b / n / 1e3:.1f} kHz")
    print(f"- Symbol duration (T = 1/Δf = N/B): {sym * 1e6:.1f} μs")
    print("\nPath Information:")
    print(f"- Maximum path delay: {max_delay * 1e6:.1f} μs")
    print(f"- Excess delay: {(max_delay - sym) * 1e6:.1f} μs")
    print("\nPaths arriving after the symbol duration will be clipped.")
    print("To avoid clipping, either:")
    print("1. Increase the number of subcarriers (N)")
    print("2. Decrease the bandwidth (B)")
    print("3. Switch to time-domain channel generation (set ch_params['freq_domain'] = 0)")
    print(bar)
    return True


# ------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("deepmimo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


@dataclass
class ChannelInfo:
    """By-products of one call: masks exactly as the reference defines them."""
    fov_mask: Optional[np.ndarray] = None      # Dataset._fov_mask (None when the reference builds none)
    valid: Optional[np.ndarray] = None         # ~isnan(power)[:, :num_paths]
    clip: Optional[np.ndarray] = None          # FD: delay_n >= N
    path_slot: Optional[np.ndarray] = None     # TD: output slot of each path column, -1 = none
    kernel: str = ""
    launches: int = 0


class ChannelPlan:
    """Device-resident inputs + packed descriptor.  `run()` launches the fused kernel for a user range.

    Holds: seven float32 [n, n_cols] path matrices, optional per-user UE rotation [n,3] float64,
    optional Doppler [n, n_cols] float32, selected subcarriers (int32) and snapshot times (float64).
    """

    def __init__(self, spec: ChannelSpec, arrays: dict, *, doppler=None, device=None):
        torch = _torch()
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.spec = spec
        ref = arrays["power"]
        self.n_users, self.n_cols = int(ref.shape[0]), int(ref.shape[1])
        if self.n_cols > MAX_COLS:
            raise ValueError(f"{self.n_cols} path columns > {MAX_COLS} supported per launch")
        # float32 is the reference's storage type (deepmimo/consts.py:65); seven float64 matrices select the all-float64
        # prologue (NumPy's dtype flow for float64 inputs); a mixture follows neither flow
        kinds = {str(getattr(arrays[k], "dtype", None)).replace("torch.", "") for k in PATH_KEYS}
        self.f64 = kinds == {"float64"}
        if not self.f64 and kinds != {"float32"}:
            raise TypeError(f"the seven path matrices must all be float32 (the reference's storage type, deepmimo/consts.py:65) "
                            f"or all float64; got {sorted(kinds)}")
        in_dtype = torch.float64 if self.f64 else torch.float32
        self.t = {}
        for k in PATH_KEYS:
            self.t[k] = self._to_dev(arrays[k], in_dtype, (self.n_users, self.n_cols), k)
        self.doppler = None if doppler is None else self._to_dev(doppler, torch.float32, (self.n_users, self.n_cols), "doppler")
        self.ue_rot = None
        if spec.ue_rot_users is not None:
            self.ue_rot = self._to_dev(spec.ue_rot_users, torch.float64, (self.n_users, 3), "ue rotation")
        self.subc = torch.as_tensor(spec.selected, dtype=torch.int32).to(self.device) if spec.freq_domain else None
        self.times = None if spec.times is None else torch.as_tensor(spec.times, dtype=torch.float64).to(self.device)
        self.desc = self._pack_desc()

    def _to_dev(self, a, dtype, shape, name):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"'{name}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
        if t.dtype != dtype:
            raise TypeError(f"'{name}' has dtype {t.dtype}, expected {dtype}")
        return t.to(self.device, non_blocking=True).contiguous()

    def _pack_desc(self) -> _lib.DmkDesc:
        s, d = self.spec, _lib.DmkDesc()
        d.bs_shape[:] = s.bs_shape
        d.ue_shape[:] = s.ue_shape
        d.bs_spacing, d.ue_spacing = s.bs_spacing, s.ue_spacing
        d.bs_rot_deg[:] = [float(v) for v in s.bs_rot]
        d.ue_rot_deg[:] = [float(v) for v in s.ue_rot]
        d.bs_fov_deg[:] = [360.0, 180.0] if s.bs_fov is None else [float(v) for v in s.bs_fov[:2]]
        d.ue_fov_deg[:] = [360.0, 180.0] if s.ue_fov is None else [float(v) for v in s.ue_fov[:2]]
        d.fov_side_enabled[:] = [int(s.fov_side[0]), int(s.fov_side[1])]
        d.fov_any = int(s.fov_any)
        d.pattern[:] = s.patterns
        d.num_paths = max(0, s.num_paths)
        d.n_cols = self.n_cols
        d.n_subcarriers = s.n_subcarriers
        d.n_selected = len(s.selected)
        d.subcarriers = self.subc.data_ptr() if (self.subc is not None and self.subc.numel() and s.subc_step == 0) else None
        d.subc_start, d.subc_step = s.subc_start, s.subc_step
        d.bandwidth = s.bandwidth
        d.rx_filter = s.rx_filter
        d.flags = _lib.FLAG_F64_INPUTS if self.f64 else 0
        d.n_times = 0 if s.times is None else len(s.times)
        d.times = None if self.times is None else self.times.data_ptr()
        return d

    # -- shapes
    def out_shape(self, n_users: Optional[int] = None) -> Tuple[int, ...]:
        return self.spec.out_shape(self.n_users if n_users is None else n_users, self.n_cols)

    def alloc_out(self, n_users: Optional[int] = None):
        torch = _torch()
        return torch.empty(self.out_shape(n_users), dtype=torch.complex64, device=self.device)

    def alloc_masks(self, n_users: Optional[int] = None):
        torch = _torch()
        n = self.n_users if n_users is None else n_users
        mk = lambda dt: torch.empty((n, self.n_cols), dtype=dt, device=self.device)
        m = {"fov": mk(torch.uint8), "valid": mk(torch.uint8)}
        if self.spec.freq_domain:
            m["clip"] = mk(torch.uint8)
        else:
            m["slot"] = mk(torch.int32)
        return m

    # -- launch
    def run(self, out, start: int = 0, stop: Optional[int] = None, masks: Optional[dict] = None, stream=None,
            independent: bool = False):
        """Launch the fused kernel for users [start, stop) into `out` (complex64 CUDA tensor whose
        first dimension is stop-start).  `masks` tensors (from alloc_masks) are indexed the same way.
        `independent=True` (DMK_FLAG_INDEPENDENT_LAUNCH, include/dmk.h) tells the library that this launch does not
        touch anything the previous kernel on the stream writes -- consecutive chunks of one user range going to
        different buffers -- so it may start while that kernel drains its tail."""
        torch = _torch()
        stop = self.n_users if stop is None else stop
        n = stop - start
        if n < 0 or start < 0 or stop > self.n_users:
            raise ValueError(f"user range [{start}, {stop}) outside [0, {self.n_users})")
        exp = self.out_shape(n)
        if tuple(out.shape) != exp or out.dtype != torch.complex64 or not out.is_contiguous() or out.device != self.device:
            raise ValueError(f"out must be a contiguous complex64 tensor of shape {exp} on {self.device}")
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        ptr = lambda t, row=start: None if t is None else t.data_ptr() + row * t.stride(0) * t.element_size()
        m = masks or {}
        mp = lambda k: None if m.get(k) is None else m[k].data_ptr()
        self.desc.flags = (_lib.FLAG_INDEPENDENT_LAUNCH if independent else 0) | (_lib.FLAG_F64_INPUTS if self.f64 else 0)
        self.desc.kernel_hint = _lib.kernel_hint_from_env("DMK_FD_KERNEL")
        self.desc.ws_helpers = int(os.environ.get("DMK_WS_HELPERS", "0") or 0)
        self.desc.ws_split = int(os.environ.get("DMK_WS_SPLIT", "0") or 0)
        common = [ctypes.byref(self.desc)] + [ptr(self.t[k]) for k in PATH_KEYS] + \
                 [ptr(self.ue_rot), ptr(self.doppler), n, self.n_cols, out.data_ptr()]
        with torch.cuda.device(self.device):
            if self.spec.freq_domain:
                rc = self.lib.dmk_channels_fd(*common, mp("fov"), mp("valid"), mp("clip"), st.cuda_stream)
            elif m.get("tau") is not None:
                rc = self.lib.dmk_channels_td_tau(*common, mp("fov"), mp("valid"), mp("slot"), mp("tau"), st.cuda_stream)
            else:
                rc = self.lib.dmk_channels_td(*common, mp("fov"), mp("valid"), mp("slot"), st.cuda_stream)
        _lib.check(rc)
        return out

    def run_beams(self, beams_dev, out, masks: Optional[dict] = None, stream=None):
        """Fused beam amplitude map (dmk_beam_amplitude_fd): `beams_dev` complex64 CUDA [n_beams, M_t], `out` float32 CUDA
        [n_users, n_beams] = mean over RX elements and subcarriers of |beams @ H|.  H is never written."""
        torch = _torch()
        nb = int(beams_dev.shape[0])
        if tuple(beams_dev.shape) != (nb, self.spec.m_tx) or beams_dev.dtype != torch.complex64 or not beams_dev.is_contiguous():
            raise ValueError(f"beams must be a contiguous complex64 CUDA tensor [n_beams, {self.spec.m_tx}]")
        if tuple(out.shape) != (self.n_users, nb) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 CUDA tensor [{self.n_users}, {nb}]")
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        m = masks or {}
        mp = lambda k: None if m.get(k) is None else m[k].data_ptr()
        self.desc.flags = _lib.FLAG_F64_INPUTS if self.f64 else 0
        self.desc.kernel_hint = _lib.kernel_hint_from_env("DMK_BF_KERNEL")
        with torch.cuda.device(self.device):
            rc = self.lib.dmk_beam_amplitude_fd(
                ctypes.byref(self.desc), *[self.t[k].data_ptr() for k in PATH_KEYS],
                None if self.ue_rot is None else self.ue_rot.data_ptr(), self.n_users, self.n_cols,
                beams_dev.data_ptr(), nb, out.data_ptr(), mp("fov"), mp("valid"), mp("clip"), st.cuda_stream)
        _lib.check(rc)
        return out

    def info_from_masks(self, masks: dict) -> ChannelInfo:
        p = self.spec.n_paths_eff(self.n_cols)
        info = ChannelInfo(kernel=_lib.last_kernel(), launches=_lib.launch_count())
        fov = masks["fov"].cpu().numpy().astype(bool)
        info.fov_mask = fov if self.spec.fov_any else None
        info.valid = masks["valid"].cpu().numpy().astype(bool)[:, :p]
        if "clip" in masks:
            info.clip = masks["clip"].cpu().numpy().astype(bool)[:, :p]
        if "slot" in masks:
            info.path_slot = masks["slot"].cpu().numpy()[:, :p]
        return info


# ------------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------------
def _dataset_arrays(dataset) -> dict:
    return {k: dataset[k] for k in PATH_KEYS}


def _get(dataset, key, default=None):
    try:
        return dataset.get(key, default)
    except AttributeError:
        try:
            return dataset[key]
        except KeyError:
            return default


C_LIGHT = 299792458          # deepmimo_v3/consts.py:112


def constant_doppler_shift(dataset, params, carrier_freq=None):
    """`enable_doppler = 1` (channel.py:50): the one Doppler definition in the reference tree is v3's
    (deepmimo_v3/generator/python/construct_deepmimo.py:267-280) -- with per-path radial velocity `doppler_vel` [m/s] and
    acceleration `doppler_acc` [m/s^2] of a dynamic scenario every frequency-domain path gain is multiplied by the constant phase
    exp(-j 2 pi f_c (v tau / c + a tau^2 / (2 c))), tau = ToA.  Returned as the per-path Doppler shift f_D [Hz] that produces
    exactly that phase at the single snapshot time t = 1 s, or None when the flag is off, the branch is time-domain (v3 applies no
    Doppler there, :90-91) or the dataset carries no velocities (v3: `doppler_available` false -> silently static)."""
    if not params.get("enable_doppler") or not params.get("freq_domain", 1):
        return None
    vel, acc = _get(dataset, "doppler_vel"), _get(dataset, "doppler_acc")
    if vel is None or acc is None:
        return None
    if int(params["ofdm"].get("rx_filter", 0)):
        raise NotImplementedError("enable_doppler with ofdm.rx_filter = 1 (per-tap Doppler of construct_deepmimo.py:274) is not implemented")
    if carrier_freq is None:
        rt = _get(dataset, "rt_params")
        carrier_freq = None if rt is None else (rt.get("frequency") if hasattr(rt, "get") else getattr(rt, "frequency", None))
    if carrier_freq is None:
        raise ValueError("enable_doppler needs the carrier frequency: dataset['rt_params']['frequency'] or carrier_freq=")
    tau = np.asarray(dataset["delay"], dtype=np.float64)
    v, a = np.asarray(vel, dtype=np.float64), np.asarray(acc, dtype=np.float64)
    return (-float(carrier_freq) * (v * tau / C_LIGHT + a * tau ** 2 / (2 * C_LIGHT))).astype(np.float32)


def make_plan(dataset, params=None, *, times=None, doppler=None, device=None, seed_numpy_rng=True,
              warn=True) -> Tuple[ChannelPlan, Any]:
    """Validate parameters against the dataset and move its path matrices to the device."""
    arrays = _dataset_arrays(dataset)
    n_ue = int(np.asarray(arrays["power"]).shape[0]) if not hasattr(arrays["power"], "device") else int(arrays["power"].shape[0])
    if params is None:
        params = _get(dataset, "ch_params")
        if params is None:
            params = ChannelGenParameters()
    if hasattr(dataset, "set_channel_params"):
        dataset.set_channel_params(params)             # validate + deep copy + cache invalidation (dataset.py:197-222)
    else:
        params.validate(n_ue)
    with _RNG_LOCK:                                    # seed + draw must not interleave between host threads (MacroDataset over devices)
        if seed_numpy_rng:
            np.random.seed(1001)                       # dataset.py:250 (global RNG side effect kept)
        spec = parse_spec(params, n_ue, bs_fov=_get(dataset, "bs_fov"), ue_fov=_get(dataset, "ue_fov"), times=times,
                          seed_numpy_rng=False)
    if doppler is None and times is not None:
        doppler = _get(dataset, "doppler")
    plan = ChannelPlan(spec, arrays, doppler=doppler if times is not None else None, device=device)
    if warn and spec.freq_domain and isinstance(arrays["delay"], np.ndarray):
        delay_overflow_warning(arrays["delay"], spec, spec.n_paths_eff(plan.n_cols))
    return plan, params


def pinned_cap_bytes() -> int:
    """Largest result `compute_channels(host_memory='auto')` allocates as page-locked memory."""
    return int(float(os.environ.get("DMK_PINNED_CAP_GIB", "16")) * (1 << 30))


def default_chunk_users(plan: ChannelPlan, budget_bytes: int = 1 << 30) -> int:
    per_user = plan.spec.coefs_per_user(plan.n_cols) * 8
    return max(1, min(plan.n_users, budget_bytes // max(per_user, 1)))


def chunk_is_independent(i: int, n_buffers: int) -> bool:
    """Whether chunk `i` of a chain of launches that cycles through `n_buffers` output buffers may carry
    DMK_FLAG_INDEPENDENT_LAUNCH (include/dmk.h, ordering contract).  A flagged launch only waits for the previous launch to
    have *begun*, so a run of s flagged launches can overlap with s predecessors; chunk i reuses the buffer of chunk
    i - n_buffers, hence at most n_buffers - 1 consecutive flagged launches, then one plain stream-ordered launch."""
    return n_buffers > 1 and (i % n_buffers) != 0


def iter_channels(plan: ChannelPlan, chunk_users: Optional[int] = None, n_buffers: int = 3) -> Iterator[Tuple[int, int, Any]]:
    """Stream H in chunks through a ring of `n_buffers` device buffers: yields (start, stop, tensor).

    The yielded tensor is valid until `n_buffers - 1` further chunks have been requested; the consumer
    must be stream-ordered after the producing stream (the current stream) or synchronise.
    """
    chunk = default_chunk_users(plan, 4 << 30) if chunk_users is None else int(chunk_users)
    ring = [plan.alloc_out(min(chunk, plan.n_users)) for _ in range(max(1, n_buffers))]
    i = 0
    for start in range(0, plan.n_users, chunk):
        stop = min(start + chunk, plan.n_users)
        buf = ring[i % len(ring)][: stop - start]
        plan.run(buf, start, stop, independent=chunk_is_independent(i, len(ring)))   # chunks of one range, different buffers
        yield start, stop, buf
        i += 1


def compute_channels(dataset, params=None, *, out: str = "numpy", device=None, chunk_users: Optional[int] = None,
                     times=None, doppler=None, return_info: bool = False, cache: bool = True, host_out=None,
                     seed_numpy_rng: bool = True, warn: bool = True, carrier_freq=None, host_memory: str = "auto"):
    """Compute MIMO channels for every user of `dataset` on the GPU.

    Same arguments, layout and caching behaviour as the reference's `Dataset.compute_channels`:
    returns complex64 `[n_ue, M_rx, M_tx, K]` (freq_domain) or `[n_ue, M_rx, M_tx, num_paths]`
    (time domain) and stores it as `dataset['channel']`.

    Extras (keyword only): `out='torch'` returns a CUDA tensor and skips the host copy (fast mode, not
    cached); `times` [T] appends a trailing snapshot axis with per-path Doppler `doppler` [n, P] Hz (or
    `dataset['doppler']`) -- row a11 of SURVEY.md; `return_info=True` also returns the masks
    (`ChannelInfo`); `host_out` is an optional preallocated (ideally pinned) complex64 tensor/array.
    `host_memory`: where a result allocated here lives -- 'pinned' (page-locked, D2H at PCIe line rate), 'pageable' (plain NumPy
    memory like the reference's; D2H staged through two pinned chunk buffers and a host copy thread) or 'auto' (pinned up to
    `pinned_cap_bytes()`, default 16 GiB or DMK_PINNED_CAP_GIB, and pageable beyond that or when page-locking fails).
    `params.enable_doppler = 1` applies v3's constant per-path Doppler phase when the dataset carries `doppler_vel` /
    `doppler_acc` (see `constant_doppler_shift`); the output keeps the reference's 4-D shape.
    """
    torch = _torch()
    if out not in ("numpy", "torch"):
        raise ValueError("out must be 'numpy' or 'torch'")
    squeeze_time = False
    if times is None and doppler is None:
        p_eff = params if params is not None else (_get(dataset, "ch_params") or ChannelGenParameters())
        doppler = constant_doppler_shift(dataset, p_eff, carrier_freq)
        if doppler is not None:
            times, squeeze_time = np.array([1.0]), True
    plan, params = make_plan(dataset, params, times=times, doppler=doppler, device=device,
                             seed_numpy_rng=seed_numpy_rng, warn=warn)
    n = plan.n_users
    masks = plan.alloc_masks() if return_info else None

    if out == "torch":
        res = plan.alloc_out()
        if n:
            plan.run(res, 0, n, masks)
        info = plan.info_from_masks(masks) if return_info else None
        if squeeze_time:
            res = res[..., 0]
        return (res, info) if return_info else res

    # ---- host output: chunked kernels on the compute stream, D2H on a copy stream, two device buffers
    shape = plan.out_shape()
    nbytes = int(np.prod(shape)) * 8
    staged = False                       # True: destination is pageable memory, D2H goes through two pinned chunk buffers
    if host_out is None:
        mode = host_memory
        if mode not in ("auto", "pinned", "pageable"):
            raise ValueError("host_memory must be 'auto', 'pinned' or 'pageable'")
        if mode == "auto":
            mode = "pinned" if nbytes <= pinned_cap_bytes() else "pageable"
        host_t = None
        if mode == "pinned":
            try:
                # torch's caching host allocator keeps page-locked blocks of dropped results: a loop that rebinds its result
                # alternates between two blocks and pins nothing after the second call
                host_t = torch.empty(shape, dtype=torch.complex64, pin_memory=True)
            except RuntimeError:
                if host_memory == "pinned":
                    raise
        if host_t is None:               # like the reference: plain pageable memory, no page-locked block stays held
            host_t = torch.from_numpy(np.empty(shape, dtype=np.complex64))
            staged = True
    else:
        host_t = host_out if isinstance(host_out, torch.Tensor) else torch.from_numpy(host_out)
        if squeeze_time and host_t.dim() == len(shape) - 1:
            host_t = host_t.unsqueeze(-1)
        if tuple(host_t.shape) != shape or host_t.dtype != torch.complex64:
            raise ValueError(f"host_out must be complex64 of shape {shape}")
        staged = not host_t.is_pinned()
    if n and nbytes:
        chunk = default_chunk_users(plan, (256 << 20) if staged else (1 << 30)) if chunk_users is None else max(1, int(chunk_users))
        compute = torch.cuda.current_stream(plan.device)
        copy = torch.cuda.Stream(device=plan.device)
        bufs = [plan.alloc_out(min(chunk, n)) for _ in range(2 if n > chunk else 1)]
        done = [None] * len(bufs)
        stage = [torch.empty(bufs[0].shape, dtype=torch.complex64, pin_memory=True) for _ in bufs] if staged else None
        landed = [None] * len(bufs)      # staged: future of the host thread that empties stage[b]
        pool = None
        if staged:
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(1)
            host_np = host_t.numpy()

        def unload(b, start, stop, ev):
            ev.synchronize()
            np.copyto(host_np[start:stop], stage[b][: stop - start].numpy())

        try:
            for i, start in enumerate(range(0, n, chunk)):
                stop = min(start + chunk, n)
                b = i % len(bufs)
                if done[b] is not None:
                    compute.wait_event(done[b])          # buffer b has been drained by the copy stream
                buf = bufs[b][: stop - start]
                sub = None if masks is None else {k: v[start:stop] for k, v in masks.items()}
                plan.run(buf, start, stop, sub, stream=compute)
                ev = torch.cuda.Event()
                ev.record(compute)
                copy.wait_event(ev)
                if staged and landed[b] is not None:
                    landed[b].result()                   # the host thread has emptied stage[b]
                with torch.cuda.stream(copy):
                    (stage[b][: stop - start] if staged else host_t[start:stop]).copy_(buf, non_blocking=True)
                    done[b] = torch.cuda.Event()
                    done[b].record(copy)
                if staged:
                    landed[b] = pool.submit(unload, b, start, stop, done[b])
            for f in landed:
                if f is not None:
                    f.result()
            copy.synchronize()
        finally:
            if pool is not None:
                pool.shutdown(wait=True)
    H = host_t.numpy() if host_out is None or isinstance(host_out, torch.Tensor) else host_out
    if squeeze_time:
        H = H[..., 0]
    info = plan.info_from_masks(masks) if return_info else None
    if cache:
        try:
            dataset["channel"] = H                       # dataset.py:266
        except Exception:  # noqa: BLE001 - a read-only mapping is fine
            pass
    return (H, info) if return_info else H
