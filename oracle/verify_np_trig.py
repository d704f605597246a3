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


def main():
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
