"""Rounding point R2 on the device: the CUDA restatement of NumPy's float32 sin/cos must equal
np.sin/np.cos bit-for-bit on the host that runs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_sincos(x):
    import torch
    from deepmimo_b200 import _lib
    lib = _lib.load()
    xd = torch.from_numpy(x).cuda()
    s = torch.empty_like(xd)
    c = torch.empty_like(xd)
    _lib.check(lib.dmk_np_sincosf(xd.data_ptr(), s.data_ptr(), c.data_ptr(), xd.numel(),
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return s.cpu().numpy(), c.cpu().numpy()


def test_device_np_sincosf_bit_exact_strided():
    hi = int(np.float32(np.pi).view(np.uint32)) + 1
    bits = np.arange(0, hi, 37, dtype=np.uint32)                      # ~29 M samples over [0, pi]
    x = bits.view(np.float32)
    s, c = _device_sincos(x)
    assert np.array_equal(s.view(np.uint32), np.sin(x).view(np.uint32))
    assert np.array_equal(c.view(np.uint32), np.cos(x).view(np.uint32))


def test_device_np_sincosf_degrees_grid_and_specials():
    rng = np.random.default_rng(7)
    deg = np.concatenate([rng.uniform(0, 180, 4_000_000), np.arange(0, 180.5, 0.5), rng.uniform(-360, 360, 500_000)])
    x = np.deg2rad(deg.astype(np.float32))
    x = np.concatenate([x, np.array([0.0, -0.0, np.nan, np.float32(np.pi), np.float32(np.pi / 2), float.fromhex('0x1.f6a7a4p+1')], np.float32)])
    s, c = _device_sincos(x)
    ok = ~np.isnan(x)
    assert np.array_equal(s.view(np.uint32)[ok], np.sin(x).view(np.uint32)[ok])
    assert np.array_equal(c.view(np.uint32)[ok], np.cos(x).view(np.uint32)[ok])
    assert np.array_equal(s[ok], np.sin(x)[ok]) and np.array_equal(c[ok], np.cos(x)[ok])
    assert np.isnan(s[~ok]).all() and np.isnan(c[~ok]).all()


def test_device_matches_c_emulation():
    import ctypes, os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "libnptrig.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_build/libnptrig.so not built")
    lib = ctypes.CDLL(so)
    x = np.random.default_rng(3).uniform(-6.3, 6.3, 2_000_000).astype(np.float32)
    o = np.empty_like(x)
    s, c = _device_sincos(x)
    lib.np_sinf_emul_array.argtypes = lib.np_cosf_emul_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.np_sinf_emul_array(x.ctypes.data, o.ctypes.data, x.size)
    assert np.array_equal(o.view(np.uint32), s.view(np.uint32))
    lib.np_cosf_emul_array(x.ctypes.data, o.ctypes.data, x.size)
    assert np.array_equal(o.view(np.uint32), c.view(np.uint32))
