"""Build libpicles_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "picles_b200", "csrc")
OUT = os.path.join(ROOT, "picles_b200", "libpicles_b200.so")
SOURCES = ["picles_kernels.cu", "picles_capi.cu", "picles1d.cu"]
HEADERS = ["physics.h", "physics1d.h", "pmath.h", "pmath_body.h", "pmath_exptab.h", "pmath_trig.h", "wind_mesh.h", "stiff.h", "pmath_dual.h", "picles_device.h", os.path.join("..", "..", "include", "picles_b200.h")]

# --fmad=false: physics.h writes every fused multiply-add explicitly so device results are
# bit-identical to the CPU oracle (see pmath.h).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "--shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=(), out: str | None = None) -> str:
    if out is None and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", out or OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose or r.stderr.strip():
        sys.stderr.write(r.stderr)
    return out or OUT


if __name__ == "__main__":
    build(force=True, verbose=True, extra=tuple(sys.argv[1:]))
