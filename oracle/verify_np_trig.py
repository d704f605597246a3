"""oracle/verify_np_trig.py -- TEST INFRASTRUCTURE.

Exhaustive check of oracle/np_trig_emul.c against np.sin / np.cos (float32) on the
host that runs the oracle.  SURVEY.md H1 asks for this to be re-verified on the
box that runs the oracle because the arithmetic is NumPy's, not the reference's.

    python oracle/verify_np_trig.py            # all float32 in [-2pi, 2pi]  (~6 min, 1 core)
    python oracle/verify_np_trig.py --quick    # 2^24 strided samples        (seconds)
"""
import ctypes
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    lib = ctypes.CDLL(os.path.join(HERE, "_build", "libnptrig.so"))
    for f in (lib.np_sinf_emul_array, lib.np_cosf_emul_array):
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        f.restype = None
    return lib


def mismatches(lib, bits):
    x = bits.view(np.float32)
    o = np.empty_like(x)
    lib.np_sinf_emul_array(x.ctypes.data, o.ctypes.data, x.size)
    bs = int((o.view(np.uint32) != np.sin(x).view(np.uint32)).sum())
    lib.np_cosf_emul_array(x.ctypes.data, o.ctypes.data, x.size)
    bc = int((o.view(np.uint32) != np.cos(x).view(np.uint32)).sum())
    return bs, bc


def unit_interval():
    """|np.sin(x)|, |np.cos(x)| <= 1 for EVERY finite float32 (all 2^32 bit patterns; ~3 min).  dmk_prologue.cuh:
    side_angles_trivial relies on it.  Result with NumPy 2.3.5: maxima 1.0 and 1.0, no violation."""
    t0 = time.time(); worst_c = worst_s = 0.0; bad = 0
    for b0 in range(0, 1 << 32, 1 << 26):
        x = np.arange(b0, b0 + (1 << 26), dtype=np.uint64).astype(np.uint32).view(np.float32)
        fin = np.isfinite(x)
        with np.errstate(invalid="ignore"):
            c = np.cos(x); s = np.sin(x)
        bad += int(((np.abs(c) > 1) | (np.abs(s) > 1) | np.isnan(c) | np.isnan(s))[fin].sum())
        worst_c = max(worst_c, float(np.abs(c[fin]).max())); worst_s = max(worst_s, float(np.abs(s[fin]).max()))
    print(f"numpy {np.__version__}: every finite float32: max |cos| {worst_c}, max |sin| {worst_s}, violations {bad}  [{time.time() - t0:.0f}s]")
    return 0 if bad == 0 else 1


def main():
    if "--unit-interval" in sys.argv:
        return unit_interval()
    quick = "--quick" in sys.argv
    lib = load()
    hi = int(np.float32(2 * np.pi).view(np.uint32)) + 1
    stride = 64 if quick else 1
    tot = bs = bc = 0
    t0 = time.time()
    for sign in (0, 0x80000000):
        for b0 in range(0, hi, 1 << 26):
            bits = np.arange(b0, min(b0 + (1 << 26), hi), stride, dtype=np.uint32) | np.uint32(sign)
            s, c = mismatches(lib, bits)
            bs += s; bc += c; tot += bits.size
    print(f"numpy {np.__version__}: checked {tot} float32 in [-2pi, 2pi] (stride {stride}): "
          f"sin mismatches {bs}, cos mismatches {bc}  [{time.time() - t0:.0f}s]")
    return 0 if (bs == 0 and bc == 0) else 1


if __name__ == "__main__":
    sys.exit(main())
