"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/*.h declares,
the ctypes mirrors match the C struct layouts, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from picles_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "picles_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(picles_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    decl = declared_symbols()
    assert decl == set(_abi.SYMBOLS), decl ^ set(_abi.SYMBOLS)
    lib = _abi.load_library()
    for name in decl:
        assert hasattr(lib, name), name
    assert lib.picles_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    prog = tmp_path / "layout.c"
    fields_p = [f for f, _ in _abi.PiclesParams._fields_]
    fields_c = [f for f, _ in _abi.PiclesCounters._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu %zu\\n", sizeof(picles_params_t), sizeof(picles_counters_t));']
    lines += [f'printf("%zu\\n", offsetof(picles_params_t, {f}));' for f in fields_p]
    lines += [f'printf("%zu\\n", offsetof(picles_counters_t, {f}));' for f in fields_c]
    lines += ["return 0;}"]
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-o", str(exe), str(prog)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == C.sizeof(_abi.PiclesParams) and int(out[1]) == C.sizeof(_abi.PiclesCounters)
    offs = [int(x) for x in out[2:]]
    exp = [getattr(_abi.PiclesParams, f).offset for f in fields_p] + [getattr(_abi.PiclesCounters, f).offset for f in fields_c]
    assert offs == exp


def test_no_cpu_fallback_create_fails_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is exercised on the CPU box")
    lib = _abi.load_library()
    h = C.c_void_p()
    rc = lib.picles_create(C.byref(h), 0)
    assert rc == -2 and not h.value                       # PICLES_ERR_CUDA
    assert b"no CPU fallback" in lib.picles_last_error(None)
    from picles_b200.engine import B200Engine
    from common import cartesian_grid, default_params
    g = cartesian_grid(8, 8)
    with pytest.raises(_abi.PiclesError, match="ERR_CUDA"):
        B200Engine(8, 8, 0, 0, g["mask"], default_params(), M_const=g["M_const"])


def test_missing_library_raises(tmp_path):
    with pytest.raises(_abi.PiclesError, match="no CPU fallback"):
        _abi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under picles_b200/ may reference it."""
    pkg = os.path.join(ROOT, "picles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cuh", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "picles_oracle" not in txt and "oracle/" not in txt.replace("the oracle/", ""), f
