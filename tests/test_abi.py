"""The C-ABI library loads and exports every symbol include/dmk.h declares; the ctypes mirror of
struct dmk_desc has the layout the C compiler gives it.  No compute calls (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmk.h")


def test_library_exports_every_declared_symbol():
    from deepmimo_b200 import _lib
    lib = _lib.load()
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b(dmk_[a-z_0-9]+)\s*\(", src))
    assert {"dmk_channels_fd", "dmk_channels_td", "dmk_path_prologue", "dmk_np_sincosf", "dmk_last_error",
            "dmk_abi_version", "dmk_launch_count", "dmk_last_kernel"} <= declared
    for name in declared:
        assert hasattr(lib, name), f"libdmk.so does not export {name}"
    assert set(_lib.SYMBOLS) == declared
    assert lib.dmk_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define DMK_ABI_VERSION (\d+)", src).group(1))
    assert lib.dmk_launch_count() >= 0 and lib.dmk_last_error() is not None


def test_ctypes_struct_matches_c_layout(tmp_path):
    from deepmimo_b200 import _lib
    fields = [f[0] for f in _lib.DmkDesc._fields_]
    prog = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{HEADER}"', "int main(void){",
            'printf("%zu\\n", sizeof(dmk_desc));']
    prog += [f'printf("%zu\\n", offsetof(dmk_desc, {f}));' for f in fields]
    prog += ["return 0;}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(c)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.DmkDesc)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.DmkDesc, f).offset == off, f


def test_error_codes_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on CPU."""
    from deepmimo_b200 import _lib
    lib = _lib.load()
    d = _lib.DmkDesc()
    d.bs_shape[:] = [8, 1]; d.ue_shape[:] = [1, 1]; d.n_cols = 25; d.num_paths = 25
    d.n_subcarriers = 64; d.n_selected = 4; d.subc_step = 1; d.bandwidth = 10e6
    d.pattern[0] = 7
    rc = lib.dmk_channels_fd(ctypes.byref(d), *([None] * 9), 0, 25, None, None, None, None, None)
    assert rc == -2 and b"pattern" in lib.dmk_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(rc)
    d.pattern[0] = 0
    d.rx_filter = 2
    rc = lib.dmk_channels_fd(ctypes.byref(d), *([None] * 9), 0, 25, None, None, None, None, None)
    assert rc == -1 and b"rx_filter" in lib.dmk_last_error()
    d.rx_filter = 0
    d.n_cols = 99
    rc = lib.dmk_channels_fd(ctypes.byref(d), *([None] * 9), 0, 25, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc)
    d.n_cols = 25
    assert lib.dmk_channels_fd(ctypes.byref(d), *([None] * 9), 0, 25, None, None, None, None, None) == 0   # n_users == 0: no-op
    assert lib.dmk_channels_td(ctypes.byref(d), *([None] * 9), 0, 10, None, None, None, None, None) == -1  # ld < n_cols


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under deepmimo_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "deepmimo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "channel_oracle" not in txt, f
    code = "import sys; import deepmimo_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
