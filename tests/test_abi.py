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
    return set(re.findall(r"\b(picles(?:1d)?_[a-z0-9_]+)\s*\(", src))


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


def test_julia_glue_ccalls_match_the_header():
    """julia/PiCLES_B200.jl cannot be executed here (no Julia in the image), so its ccall sites are
    checked statically: every symbol exists in the C ABI and is called with as many arguments as
    the header declares, with pointer/scalar kinds in the right places; the Julia mirrors of the two
    structs list the header's fields in order."""
    src = open(os.path.join(ROOT, "julia", "PiCLES_B200.jl"), encoding="utf-8").read()
    calls = re.findall(r"ccall\(\(:(picles(?:1d)?_\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", src, flags=re.S)
    assert len(calls) >= 12
    seen = set()
    for name, ret, argt in calls:
        assert name in _abi.SYMBOLS, name
        res, args = _abi.SYMBOLS[name]
        jl = [a.strip() for a in argt.split(",") if a.strip()]
        assert len(jl) == len(args), f"{name}: Julia passes {len(jl)} arguments, the header declares {len(args)}"
        for k, (ja, ca) in enumerate(zip(jl, args)):
            is_ptr_c = ca in (_abi._vp, C.c_char_p) or hasattr(ca, "contents") or isinstance(ca, type(C.POINTER(C.c_int)))
            is_ptr_jl = ja.startswith(("Ptr{", "Ref{")) or ja == "Cstring"
            if ca is C.c_double:
                assert ja == "Cdouble", (name, k, ja)
            elif ca is C.c_int:
                assert ja == "Cint", (name, k, ja)
            else:
                assert is_ptr_jl and is_ptr_c, (name, k, ja, ca)
        assert ret == ("Cstring" if res is C.c_char_p else "Cint"), name
        seen.add(name)
    for must in ("picles_create", "picles_set_grid", "picles_set_grid_metric", "picles_set_params", "picles_seed",
                 "picles_step", "picles_step_strip", "picles_get_state", "picles_get_counters", "picles_comm_init",
                 "picles_set_wind_midlevels", "picles_set_wind_mesh", "picles_step_wind_mesh",
                 "picles1d_create", "picles1d_set_grid", "picles1d_set_params", "picles1d_seed", "picles1d_step", "picles1d_get_state"):
        assert must in seen, must
    # struct mirrors: same field names in the same order as the ctypes mirrors (checked against the header above)
    for jl_name, ct in (("PiclesParams", _abi.PiclesParams), ("PiclesCounters", _abi.PiclesCounters)):
        body = re.search(r"struct %s\b(.*?)\nend" % jl_name, src, flags=re.S).group(1)
        body = re.sub(r"#.*", "", body)
        fields = re.findall(r"(\w+)::", body)
        assert fields == [f for f, *_ in ct._fields_], jl_name
