"""oracle/np_trig_emul.c against NumPy's float32 sin/cos on this host (rounding point R2).
The exhaustive run over all float32 in [-2pi, 2pi] is oracle/verify_np_trig.py (minutes); this is the
strided version that fits the CPU suite."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libnptrig.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(SO)
    for f in (lib.np_sinf_emul_array, lib.np_cosf_emul_array):
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    return lib


def _both(lib, x):
    s = np.empty_like(x); c = np.empty_like(x)
    lib.np_sinf_emul_array(x.ctypes.data, s.ctypes.data, x.size)
    lib.np_cosf_emul_array(x.ctypes.data, c.ctypes.data, x.size)
    return s, c


def test_strided_sweep_bit_exact(lib):
    hi = int(np.float32(2 * np.pi).view(np.uint32)) + 1
    for sign in (0, 0x80000000):
        bits = np.arange(0, hi, 251, dtype=np.uint32) | np.uint32(sign)
        x = bits.view(np.float32)
        s, c = _both(lib, x)
        assert np.array_equal(s.view(np.uint32), np.sin(x).view(np.uint32))
        assert np.array_equal(c.view(np.uint32), np.cos(x).view(np.uint32))


def test_degree_grid_and_fused_quadrant_case(lib):
    deg = np.concatenate([np.arange(0, 180.25, 0.25), np.random.default_rng(5).uniform(0, 180, 1_000_000)]).astype(np.float32)
    x = np.deg2rad(deg)
    x = np.concatenate([x, np.array([float.fromhex("0x1.f6a7a4p+1")], np.float32)])   # needs the fused q = fma(x, 2/pi, magic)
    s, c = _both(lib, x)
    assert np.array_equal(s.view(np.uint32), np.sin(x).view(np.uint32))
    assert np.array_equal(c.view(np.uint32), np.cos(x).view(np.uint32))
    s, c = _both(lib, np.array([np.nan], np.float32))
    assert np.isnan(s[0]) and np.isnan(c[0])


def test_float32_sin_cos_stay_inside_unit_interval():
    """dmk_prologue.cuh: side_angles_trivial relies on |cos(x)| <= 1 for NumPy's float32 cos (then arccos of it is never NaN for a
    finite angle).  Strided sweep over every finite float32 here; the exhaustive sweep (all 2^32 bit patterns, ~1 min) is
    `python oracle/verify_np_trig.py --unit-interval`."""
    bits = np.arange(0, 1 << 32, 4099, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    x = x[np.isfinite(x)]
    assert np.abs(np.cos(x)).max() <= 1.0 and np.abs(np.sin(x)).max() <= 1.0
