"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of the host-side kernel emulator (tests/emu/rrx_emu.cc)."""
from __future__ import annotations

import ctypes as C
import functools
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# RRX_EMU_DEFS="A B=1": extra #defines for the emulator build (the same switches RRX_JIT_DEFS gives the kernels)
EMU_DEFS = os.environ.get("RRX_EMU_DEFS", "").split()
LIB = os.path.join(HERE, "emu", "_build", "librrx_emu" + "".join("_" + d.replace("=", "-") for d in EMU_DEFS) + ".so")
SRCS = [os.path.join(HERE, "emu", "rrx_emu.cc"), os.path.join(ROOT, "roaringregex_b200", "csrc", "rrx_compile.cc")]
DEPS = SRCS + [os.path.join(ROOT, "roaringregex_b200", "csrc", f) for f in ("rrx_tile.cuh", "rrx_compile.h", "rrx_prog.h")]


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas"] + ["-D" + d for d in EMU_DEFS] + ["-o", LIB] + SRCS, check=True)


@functools.lru_cache(maxsize=None)
def _lib():
    build()
    lib = C.CDLL(LIB)
    lib.emu_scan.restype = C.c_int
    lib.emu_scan.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_uint32]
    lib.emu_pending_size.restype = C.c_uint32
    lib.emu_set_sparse.argtypes = [C.c_int]
    lib.emu_sparse_stats.argtypes = [C.c_void_p, C.c_int]
    return lib


def set_sparse(mode: int):
    """0 dense only, 1 candidate-driven with dense fallback (default, = the kernel), 2 run both and compare."""
    _lib().emu_set_sparse(mode)


def sparse_stats(reset=True):
    out = np.zeros(3, dtype=np.uint64)
    _lib().emu_sparse_stats(out.ctypes.data, int(reset))
    global last_fast_tiles
    last_fast_tiles = int(out[2])
    return int(out[0]), int(out[1])


last_fast_tiles = 0


def scan(pattern: str, text, reduce=True, fast=True, left_ctx=0, base=0, entry=None, cap=None, donly=True):
    """Returns dict(matches (M,2) int64, n, pending, err, shard_comp, info)."""
    buf = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text)
    # keep 16 spare bytes around so that &buf[0] is valid for empty inputs
    padded = np.concatenate([buf, np.zeros(16, np.uint8)])
    n = len(buf) - left_ctx
    cap = cap if cap is not None else max(64, n + 1)
    out = np.full((cap, 2), -7, dtype=np.int64)
    counts = np.zeros(2, dtype=np.uint64)
    err = np.zeros(1, dtype=np.uint32)
    comp = np.zeros(1 + 3 * 32, dtype=np.uint32)
    info = np.zeros(12, dtype=np.uint32)
    ent = None
    if entry is not None:
        ent = np.zeros(32, dtype=np.uint32)
        ent[:len(entry)] = entry
    psz = _lib().emu_pending_size()
    pend = np.zeros(64 * psz, dtype=np.uint8)
    rc = _lib().emu_scan(pattern.encode("latin-1"), int(bool(reduce)) | (0 if donly else 2), int(fast), padded.ctypes.data, left_ctx, n, base,
                         ent.ctypes.data if ent is not None else None, out.ctypes.data, cap, counts.ctypes.data,
                         err.ctypes.data, comp.ctypes.data, info.ctypes.data, pend.ctypes.data, 64)
    if rc:
        raise ValueError(f"emu_scan rc={rc}")
    m = int(counts[0])
    return dict(matches=out[:min(m, cap)].copy(), n=m, pending=int(counts[1]), err=int(err[0]) & ~16, soft=int(err[0]) & 16, shard_comp=comp,
                info=dict(NW=int(info[0]), P=int(info[1]), n_terms=int(info[2]), n_bicl=int(info[3]),
                          use_rows=int(info[4]), static_depth=int(info[5]), sparse_ok=int(info[6]), cand_byte=int(info[7]), donly=int(info[8])), pend_raw=pend, pend_size=psz)
