"""Build recipe for librrx.so (nvcc, sm_100a only) — in-tree, so the .so travels with gpurun."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
_VARIANT = os.environ.get("RRX_VARIANT", "")          # e.g. "timing": builds librrx_timing.so with -DRRX_PHASE_TIMING
LIB = os.path.join(PKG, f"librrx_{_VARIANT}.so" if _VARIANT else "librrx.so")
SOURCES = ["rrx_capi.cu", "rrx_scan.cu", "rrx_order.cu", "rrx_corpus.cu", "rrx_compile.cc"]
HEADERS = ["rrx_compile.h", "rrx_prog.h", "rrx_scan.h", "rrx_tile.cuh", os.path.join("..", "..", "include", "rrx.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xcompiler", "-Wno-unknown-pragmas"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DRRX_PHASE_TIMING"] if _VARIANT == "timing" else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building librrx.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
