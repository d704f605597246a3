"""ctypes binding of libdmk.so (include/dmk.h).  There is no CPU fallback: if the library cannot be
loaded this module raises, and every compute entry point of the package fails loudly."""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

c_f32p = ctypes.c_void_p
_lock = threading.Lock()
_lib = None


class DmkDesc(ctypes.Structure):
    """struct dmk_desc (include/dmk.h) -- field order and types must match the header."""
    _fields_ = [
        ("bs_shape", ctypes.c_int32 * 2),
        ("ue_shape", ctypes.c_int32 * 2),
        ("bs_spacing", ctypes.c_double),
        ("ue_spacing", ctypes.c_double),
        ("bs_rot_deg", ctypes.c_double * 3),
        ("ue_rot_deg", ctypes.c_double * 3),
        ("bs_fov_deg", ctypes.c_double * 2),
        ("ue_fov_deg", ctypes.c_double * 2),
        ("fov_side_enabled", ctypes.c_int32 * 2),
        ("fov_any", ctypes.c_int32),
        ("pattern", ctypes.c_int32 * 2),
        ("num_paths", ctypes.c_int32),
        ("n_cols", ctypes.c_int32),
        ("n_subcarriers", ctypes.c_int32),
        ("n_selected", ctypes.c_int32),
        ("subcarriers", ctypes.c_void_p),
        ("subc_start", ctypes.c_int32),
        ("subc_step", ctypes.c_int32),
        ("bandwidth", ctypes.c_double),
        ("rx_filter", ctypes.c_int32),
        ("n_times", ctypes.c_int32),
        ("times", ctypes.c_void_p),
        ("flags", ctypes.c_int32),
        ("kernel_hint", ctypes.c_int32),
        ("ws_helpers", ctypes.c_int32),
        ("ws_split", ctypes.c_int32),
    ]


ABI_VERSION = 3
FLAG_INDEPENDENT_LAUNCH = 1
FLAG_F64_INPUTS = 2
KERNEL_HINTS = {"": 0, "auto": 0, "fast": 0, "tile": 1, "ffma": 2, "tc": 3, "tc1": 4, "small": 5, "small1": 6, "mma": 7, "rows": 8}      # enum dmk_kernel_hint ("fast" = the default route)


def kernel_hint_from_env(var: str = "DMK_FD_KERNEL") -> int:
    """Test / A-B knob: DMK_FD_KERNEL (or DMK_BF_KERNEL) = tile | ffma | tc | tc1 | small -> dmk_desc.kernel_hint."""
    v = os.environ.get(var, "").strip().lower()
    if v not in KERNEL_HINTS:
        raise ValueError(f"{var}={v!r}: expected one of {sorted(k for k in KERNEL_HINTS if k)}")
    return KERNEL_HINTS[v]
SYMBOLS = ("dmk_channels_fd", "dmk_channels_td", "dmk_channels_td_tau", "dmk_beam_amplitude_fd", "dmk_path_prologue", "dmk_user_byproducts", "dmk_np_sincosf",
           "dmk_last_error", "dmk_abi_version", "dmk_launch_count", "dmk_last_kernel")


class DmkError(RuntimeError):
    pass


def _declare(lib: ctypes.CDLL) -> None:
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    desc_p = ctypes.POINTER(DmkDesc)
    common = [desc_p] + [vp] * 7 + [vp, vp, i64, i32, vp]
    lib.dmk_channels_fd.argtypes = common + [vp, vp, vp, vp]
    lib.dmk_channels_fd.restype = ctypes.c_int
    lib.dmk_channels_td.argtypes = common + [vp, vp, vp, vp]
    lib.dmk_channels_td.restype = ctypes.c_int
    lib.dmk_channels_td_tau.argtypes = common + [vp, vp, vp, vp, vp]
    lib.dmk_channels_td_tau.restype = ctypes.c_int
    lib.dmk_beam_amplitude_fd.argtypes = [desc_p] + [vp] * 7 + [vp, i64, i32, vp, i32, vp, vp, vp, vp, vp]
    lib.dmk_beam_amplitude_fd.restype = ctypes.c_int
    lib.dmk_path_prologue.argtypes = [desc_p] + [vp] * 5 + [vp, i64, i32, vp, vp, vp, vp]
    lib.dmk_path_prologue.restype = ctypes.c_int
    lib.dmk_user_byproducts.argtypes = [desc_p] + [vp] * 6 + [vp, vp, i64, i32, vp, vp, vp, vp, vp]
    lib.dmk_user_byproducts.restype = ctypes.c_int
    lib.dmk_np_sincosf.argtypes = [vp, vp, vp, i64, vp]
    lib.dmk_np_sincosf.restype = ctypes.c_int
    lib.dmk_last_error.restype = ctypes.c_char_p
    lib.dmk_last_kernel.restype = ctypes.c_char_p
    lib.dmk_abi_version.restype = ctypes.c_int
    lib.dmk_launch_count.restype = ctypes.c_int64


def library_path() -> str:
    return _build.LIB


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc exists) deepmimo_b200/libdmk.so."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("DMK_LIB_PATH") or _build.LIB        # DMK_LIB_PATH: an alternative build of the library (A/B timing)
        if path == _build.LIB and (not os.path.exists(path) or _build.is_stale()):
            try:
                _build.build_lib()
            except Exception as e:  # noqa: BLE001
                if not os.path.exists(path):
                    raise DmkError(f"libdmk.so is missing and could not be built ({e}); "
                                   "deepmimo_b200 has no CPU fallback") from e
                # an older library exists (e.g. a box without nvcc): say so instead of silently running stale kernels;
                # an ABI mismatch still raises below
                import warnings
                warnings.warn(f"libdmk.so is older than its sources and the rebuild failed ({str(e)[:200]}); "
                              "using the existing library", RuntimeWarning, stacklevel=2)
        lib = ctypes.CDLL(path)
        for s in SYMBOLS:
            if not hasattr(lib, s):
                raise DmkError(f"{path} does not export {s}")
        _declare(lib)
        if lib.dmk_abi_version() != ABI_VERSION:
            raise DmkError(f"libdmk ABI {lib.dmk_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().dmk_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise DmkError(f"libdmk error {rc}: {msg}")


def launch_count() -> int:
    return int(load().dmk_launch_count())


def last_kernel() -> str:
    return load().dmk_last_kernel().decode()
