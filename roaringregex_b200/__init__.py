"""roaringregex_b200 — B200-native drop-in for RoaringRegex's hot path (bit-parallel NFA scan ->
lazy-match enumeration).

Python mirror of the reference's public interface (`/root/reference/src/inc/regex.h:100-126,212-228`)
on top of the C ABI in `include/rrx.h` (librrx.so, hand-written sm_100a CUDA):

    RRegex(pattern)                 regex.h:224 / Parser.cpp:161   (raises InvalidExpression like the
                                    reference's std::runtime_error)
    RRegex.get_lazy_iter(text)      the lazy iterator the reference's README promises and
                                    src/Iterator.cpp leaves empty
    RRegex.get_acceptance_iter(text) regex.h:225 (whole-string acceptance, answered from the match list)
    it.advance() / it.value()       IteratorWrapper::operator++(int) / operator*()  (regex.h:118-119)
    Match(start, end).str()         regex.h:100-105

There is NO CPU scanning path in this package: without the CUDA extension or without a GPU the
scan calls raise.  (The CPU oracle lives under /oracle and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass

import numpy as np

from . import _build

__all__ = ["RRegex", "Match", "InvalidExpression", "RrxError", "Scanner", "lib", "corpus_iid", "corpus_records", "corpus_lines", "corpus_device", "NcclComm", "ShardedScanner"]

RRX_OK, RRX_ERR_PARSE, RRX_ERR_CUDA, RRX_ERR_UNSUPPORTED, RRX_ERR_ARG = range(5)
MODE_LAZY_V1, MODE_NO_OVERLAP = 0, 1
COMPILE_DEFAULT, COMPILE_NO_REDUCE, COMPILE_NO_DONLY = 0, 1, 2
SCAN_NO_DONLY = 1
SOFT_VERIFY_EDGE = 16


class RrxError(RuntimeError):
    pass


class InvalidExpression(RrxError):
    """The reference's `throw std::runtime_error("invalid expression")` (Parser.cpp:36,155)."""


class _Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_raw", "ref_width_class", "n_flat", "n_device", "words", "nullable",
                                          "n_terms", "n_bicliques", "uses_rows", "static_closed", "sparse_ok", "cand_byte", "donly", "carried_words")]


class _Opts(C.Structure):
    _fields_ = [("left_context", C.c_uint64), ("entry", C.c_void_p), ("mode", C.c_int), ("flags", C.c_uint32)]


class _Result(C.Structure):
    _fields_ = [("n_matches", C.c_uint64), ("capacity", C.c_uint64), ("n_pending", C.c_uint64),
                ("error_bits", C.c_uint32), ("launches", C.c_uint32), ("d_matches", C.c_void_p),
                ("soft_bits", C.c_uint32), ("reserved", C.c_uint32)]


class _ShardedResult(C.Structure):
    _fields_ = [("n_matches", C.c_uint64), ("d_matches", C.c_void_p), ("rescans", C.c_uint32), ("patched_starts", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32), ("scanner", C.c_void_p)]


_LIB = None


def lib() -> C.CDLL:
    """Load librrx.so (building it in-tree with nvcc if the sources are newer).  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.needs_build():
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            _build.build()
        elif not os.path.exists(path):
            raise RrxError("librrx.so is missing and nvcc is not available to build it")
    L = C.CDLL(path)
    vp, u64, u32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64
    sigs = {
        "rrx_compile": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
        "rrx_compile_ex": (C.c_int, [C.c_char_p, u32, C.POINTER(vp)]),
        "rrx_program_free": (None, [vp]),
        "rrx_program_info": (None, [vp, C.POINTER(_Info)]),
        "rrx_program_flat_table": (i64, [vp, C.c_int, vp, u64]),
        "rrx_program_blob": (vp, [vp, C.POINTER(u64)]),
        "rrx_program_raw_row": (i64, [vp, u32, u32, vp, u64]),
        "rrx_program_raw_finals": (i64, [vp, vp, u64]),
        "rrx_scan_device": (C.c_int, [vp, vp, u64, u64, C.c_int, C.POINTER(vp), vp]),
        "rrx_scan_host": (C.c_int, [vp, vp, u64, u64, C.c_int, C.POINTER(vp)]),
        "rrx_accepts_host": (C.c_int, [vp, vp, u64, C.POINTER(C.c_int)]),
        "rrx_matches_count": (u64, [vp]),
        "rrx_matches_data": (vp, [vp]),
        "rrx_matches_device_data": (vp, [vp]),
        "rrx_matches_free": (None, [vp]),
        "rrx_scanner_create": (C.c_int, [vp, u64, u64, C.POINTER(vp)]),
        "rrx_scanner_create_ex": (C.c_int, [vp, u64, u64, u32, C.POINTER(vp)]),
        "rrx_scanner_is_jit": (C.c_int, [vp]),
        "rrx_scanner_used_fallback": (C.c_int, [vp]),
        "rrx_scanner_kernel_variant": (C.c_int, [vp]),
        "rrx_program_jit_source": (u64, [vp, vp, u64]),
        "rrx_program_jit_compile": (C.c_int, [vp, C.c_char_p]),
        "rrx_scanner_free": (None, [vp]),
        "rrx_scanner_scan": (C.c_int, [vp, vp, u64, u64, C.POINTER(_Opts), vp]),
        "rrx_scanner_result": (C.c_int, [vp, C.POINTER(_Result), vp]),
        "rrx_scanner_shard_summary": (C.c_int, [vp, vp, u64, vp]),
        "rrx_scanner_enable_timing": (C.c_int, [vp, C.c_int]),
        "rrx_scanner_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "rrx_scanner_message_device": (C.c_int, [vp, vp, u64, vp]),
        "rrx_scanner_copy_matches": (C.c_int, [vp, vp, u64, vp]),
        "rrx_scanner_copy_pending": (C.c_int, [vp, vp, u64, C.POINTER(u64), vp]),
        "rrx_fold_entry": (C.c_int, [vp, vp, u32, u32, vp, C.POINTER(C.c_int)]),
        "rrx_fold_entries": (C.c_int, [vp, vp, u32, vp, vp]),
        "rrx_apply_summary": (C.c_int, [vp, vp, vp, vp]),
        "rrx_continue_starts": (C.c_int, [vp, vp, u64, u64, u64, vp, u64, vp, vp]),
        "rrx_sharded_create": (C.c_int, [vp, u64, u64, vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "rrx_sharded_begin": (C.c_int, [vp, vp, u64, u64, u64, vp]),
        "rrx_sharded_finish": (C.c_int, [vp, C.POINTER(_ShardedResult)]),
        "rrx_scan_sharded": (C.c_int, [vp, vp, u64, u64, u64, vp, C.POINTER(_ShardedResult)]),
        "rrx_sharded_enable_timing": (C.c_int, [vp, C.c_int]),
        "rrx_sharded_kernel_ms": (C.c_float, [vp]),
        "rrx_sharded_free": (None, [vp]),
        "rrx_digest_device": (C.c_int, [vp, u64, i64, vp, vp]),
        "rrx_scanner_digest": (C.c_int, [vp, i64, vp, vp]),
        "rrx_scanner_no_overlap": (C.c_int, [vp, C.POINTER(i64), C.POINTER(vp), vp]),
        "rrx_corpus_iid_device": (C.c_int, [vp, u64, u64, u64, vp, vp]),
        "rrx_corpus_records_device": (C.c_int, [vp, u64, u64, u64, vp]),
        "rrx_corpus_lines_device": (C.c_int, [vp, u64, u64, u64, vp]),
        "rrx_last_error": (C.c_char_p, []),
        "rrx_version": (C.c_char_p, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)          # AttributeError = a symbol include/rrx.h declares is missing
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L


def _check(rc: int):
    if rc == RRX_OK:
        return
    msg = lib().rrx_last_error().decode("latin-1")
    if rc == RRX_ERR_PARSE:
        raise InvalidExpression(msg)
    raise RrxError(f"rrx error {rc}: {msg}")


@dataclass
class Match:
    """regex.h:100-105 — half-open [start,end) into the scanned text."""
    start: int
    end: int
    _text: bytes | None = None

    def str(self) -> bytes:
        return self._text[self.start:self.end] if self._text is not None else b""


class RRegex:
    """Compiled regex (regex.h:212-228).  Owns the rrx_program handle."""

    def __init__(self, pattern, flags: int = COMPILE_DEFAULT):
        self.pattern = pattern if isinstance(pattern, bytes) else pattern.encode("latin-1")
        h = C.c_void_p()
        _check(lib().rrx_compile_ex(self.pattern, flags, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _LIB is not None:
            _LIB.rrx_program_free(h)

    # -- introspection -------------------------------------------------------------------------
    @property
    def info(self) -> dict:
        cached = getattr(self, "_info", None)           # the program is immutable
        if cached is None:
            i = _Info()
            lib().rrx_program_info(self._h, C.byref(i))
            cached = self._info = {n: int(getattr(i, n)) for n, _ in _Info._fields_}
        return dict(cached)

    def flat_table(self, which: int) -> np.ndarray:
        need = -lib().rrx_program_flat_table(self._h, which, None, 0)
        out = np.zeros(max(int(need), 1), dtype=np.uint32)
        got = lib().rrx_program_flat_table(self._h, which, out.ctypes.data, len(out))
        return out[:max(int(got), 0)]

    def blob(self) -> np.ndarray:
        n = C.c_uint64(0)
        p = lib().rrx_program_blob(self._h, C.byref(n))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(int(n.value),)).copy()

    def jit_source(self) -> str:
        n = int(lib().rrx_program_jit_source(self._h, None, 0))
        buf = C.create_string_buffer(n)
        lib().rrx_program_jit_source(self._h, buf, n)
        return buf.value.decode("latin-1")

    def jit_compile(self, path: str | None = None) -> None:
        """NVRTC-compile the specialised kernel for sm_100a (works without a GPU)."""
        _check(lib().rrx_program_jit_compile(self._h, path.encode() if path else None))

    def raw_row(self, state: int, byte: int) -> list[int]:
        w = (self.info["n_raw"] + 31) // 32
        out = np.zeros(w, dtype=np.uint32)
        lib().rrx_program_raw_row(self._h, state, byte, out.ctypes.data, w)
        return _bits_of(out)

    def raw_finals(self) -> list[int]:
        w = (self.info["n_raw"] + 31) // 32
        out = np.zeros(w, dtype=np.uint32)
        lib().rrx_program_raw_finals(self._h, out.ctypes.data, w)
        return _bits_of(out)

    # -- scanning (GPU only) -------------------------------------------------------------------------
    def scan_host(self, text, base_offset: int = 0, mode: int = MODE_LAZY_V1) -> np.ndarray:
        """All matches of a host buffer as an (M,2) int64 array; H2D copy + kernels + D2H copy."""
        buf = _as_u8(text)
        h = C.c_void_p()
        _check(lib().rrx_scan_host(self._h, buf.ctypes.data if len(buf) else None, len(buf), base_offset, mode, C.byref(h)))
        return _take_matches(h)

    def scan_device(self, d_ptr: int, length: int, base_offset: int = 0, mode: int = MODE_LAZY_V1, stream: int = 0) -> np.ndarray:
        h = C.c_void_p()
        _check(lib().rrx_scan_device(self._h, d_ptr, length, base_offset, mode, C.byref(h), stream))
        return _take_matches(h)

    def accepts(self, text) -> bool:
        """Whole-string acceptance (the reference's AcceptanceIterator); runs on the GPU."""
        buf = _as_u8(text)
        acc = C.c_int(0)
        _check(lib().rrx_accepts_host(self._h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(acc)))
        return bool(acc.value)

    def get_lazy_iter(self, text) -> "LazyIterator":
        return LazyIterator(self, text)

    def get_acceptance_iter(self, text) -> "AcceptanceIterator":
        return AcceptanceIterator(self, text)


def _take_matches(h) -> np.ndarray:
    """(M,2) int64 view of the match list owned by the rrx_matches handle (for host scans: pinned
    memory the D2H copy landed in — no second copy); the handle is released with the array."""
    L = lib()
    n = int(L.rrx_matches_count(h))
    if n == 0:
        L.rrx_matches_free(h)
        return np.zeros((0, 2), dtype=np.int64)
    p = L.rrx_matches_data(h)
    if not p:
        L.rrx_matches_free(h)
        _check(RRX_ERR_CUDA)
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(n, 2))
    fin = weakref.finalize(arr, L.rrx_matches_free, h)
    fin.atexit = False
    return arr


def _bits_of(words: np.ndarray) -> list[int]:
    out = []
    for i, w in enumerate(words.tolist()):
        while w:
            low = w & -w
            out.append(i * 32 + low.bit_length() - 1)
            w ^= low
    return out


def _as_u8(text) -> np.ndarray:
    if isinstance(text, np.ndarray):
        assert text.dtype == np.uint8
        return np.ascontiguousarray(text)
    if isinstance(text, str):
        text = text.encode("latin-1")
    return np.frombuffer(bytes(text), dtype=np.uint8) if len(text) else np.zeros(0, np.uint8)


class LazyIterator:
    """`++` advances to the next LAZY-v1 match, `*` returns it or None when exhausted — the
    IteratorWrapper protocol of regex.h:113-122.  The whole list is produced by one GPU scan on
    first use; text is cut at the first NUL like the reference's `char*` driver (regex.h:157)."""

    def __init__(self, regex: RRegex, text):
        raw = bytes(text) if not isinstance(text, str) else text.encode("latin-1")
        nul = raw.find(b"\0")
        self._text = raw if nul < 0 else raw[:nul]
        self._regex = regex
        self._matches = None
        self._pos = -1

    def _ensure(self):
        if self._matches is None:
            self._matches = self._regex.scan_host(self._text)

    def advance(self) -> "LazyIterator":          # operator++(int)
        self._ensure()
        self._pos += 1
        return self

    def value(self):                              # operator*()
        self._ensure()
        if 0 <= self._pos < len(self._matches):
            s, e = self._matches[self._pos]
            return Match(int(s), int(e), self._text)
        return None

    def create_copy(self) -> "LazyIterator":
        c = LazyIterator.__new__(LazyIterator)
        c.__dict__.update(self.__dict__)
        return c

    def __iter__(self):
        self._ensure()
        for s, e in self._matches:
            yield Match(int(s), int(e), self._text)


class AcceptanceIterator:
    """regex.h:150-165: `++` consumes the whole NUL-terminated text; `*` is Match(0,len) iff the
    whole string is in the language."""

    def __init__(self, regex: RRegex, text):
        raw = bytes(text) if not isinstance(text, str) else text.encode("latin-1")
        nul = raw.find(b"\0")
        self._text = raw if nul < 0 else raw[:nul]
        self._regex = regex
        self._accepted = None

    def advance(self):
        self._accepted = self._regex.accepts(self._text)
        return self

    def value(self):
        if self._accepted is None:           # the reference's processor before any byte: initial state only
            return Match(0, 0, self._text) if self._regex.info["nullable"] else None
        return Match(0, len(self._text), self._text) if self._accepted else None

    def create_copy(self):
        c = AcceptanceIterator.__new__(AcceptanceIterator)
        c.__dict__.update(self.__dict__)
        return c


class Scanner:
    """Reusable device-side scan context (rrx_scanner): no allocation inside a scan call."""

    def __init__(self, regex: RRegex, max_len: int, capacity: int, jit=None):
        """jit: None = default policy (specialise for long inputs), True / False = force / forbid."""
        self.regex = regex
        h = C.c_void_p()
        flags = 0 if jit is None else (1 if jit else 2)
        _check(lib().rrx_scanner_create_ex(regex._h, max_len, capacity, flags, C.byref(h)))
        self._h = h
        self.capacity = capacity
        self._entry_keep = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _LIB is not None:
            _LIB.rrx_scanner_free(h)

    @property
    def is_jit(self) -> bool:
        return bool(lib().rrx_scanner_is_jit(self._h))

    @property
    def kernel_variant(self) -> int:
        """Specialised kernel the next scan uses: 0 lean, 1 straddle (rrx_scanner_kernel_variant)."""
        return int(lib().rrx_scanner_kernel_variant(self._h))

    @property
    def used_fallback(self) -> bool:
        """The last scan ran on the fallback (row-bounded) scanner of a thread-owned program."""
        return bool(lib().rrx_scanner_used_fallback(self._h))

    def scan(self, d_ptr: int, length: int, base_offset: int = 0, left_context: int = 0, entry=None, stream: int = 0,
             no_donly: bool = False):
        o = _Opts(left_context, None, MODE_LAZY_V1, SCAN_NO_DONLY if no_donly else 0)
        if entry is not None:
            self._entry_keep = np.ascontiguousarray(entry, dtype=np.uint32)
            o.entry = self._entry_keep.ctypes.data
        _check(lib().rrx_scanner_scan(self._h, d_ptr, length, base_offset, C.byref(o), stream))

    def result(self, stream: int = 0) -> dict:
        r = _Result()
        _check(lib().rrx_scanner_result(self._h, C.byref(r), stream))
        return dict(n_matches=int(r.n_matches), capacity=int(r.capacity), n_pending=int(r.n_pending),
                    error_bits=int(r.error_bits), launches=int(r.launches), d_matches=int(r.d_matches or 0),
                    soft_bits=int(r.soft_bits))

    def matches(self, stream: int = 0) -> np.ndarray:
        """Match list of the last scan as an (M,2) int64 host array."""
        r = self.result(stream)
        n = min(r["n_matches"], r["capacity"])
        out = np.empty((n, 2), dtype=np.int64)
        _check(lib().rrx_scanner_copy_matches(self._h, out.ctypes.data, n, stream))
        return out

    def digest(self, end_limit: int = (1 << 62), stream: int = 0) -> tuple:
        """(count, sum, xor, order_violations) of the last scan's device match list (rrx_scanner_digest)."""
        d = np.zeros(4, dtype=np.uint64)
        _check(lib().rrx_scanner_digest(self._h, end_limit, d.ctypes.data, stream))
        return tuple(int(x) for x in d)

    def no_overlap(self, carry: int = 0, stream: int = 0):
        """Non-overlapping view of the last scan's list, chained by `carry` (rrx_scanner_no_overlap).
        -> ((M,2) int64 host array, new carry)"""
        c = C.c_int64(carry)
        h = C.c_void_p()
        _check(lib().rrx_scanner_no_overlap(self._h, C.byref(c), C.byref(h), stream))
        return _take_matches(h), int(c.value)

    def enable_timing(self, on: bool = True):
        _check(lib().rrx_scanner_enable_timing(self._h, int(on)))

    def kernel_ms(self) -> float:
        """Duration of the last scan kernel launch alone (CUDA events on the launching stream)."""
        ms = C.c_float(0)
        _check(lib().rrx_scanner_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def message_device(self, d_ptr: int, cap_words: int, stream: int = 0):
        """Enqueue a device-to-device copy of [shard composite | counters] (rrx_scanner_message_device)."""
        _check(lib().rrx_scanner_message_device(self._h, d_ptr, cap_words, stream))

    def shard_summary(self, stream: int = 0) -> np.ndarray:
        w = self.regex.info["words"]
        out = np.zeros(1 + 3 * w, dtype=np.uint32)
        _check(lib().rrx_scanner_shard_summary(self._h, out.ctypes.data, len(out), stream))
        return out


def corpus_iid(d_ptr: int, offset: int, length: int, seed: int, alpha: bytes, stream: int = 0):
    a = np.frombuffer(bytes(alpha), dtype=np.uint8)
    assert len(a) == 256
    _check(lib().rrx_corpus_iid_device(d_ptr, offset, length, seed, a.ctypes.data, stream))


def corpus_records(d_ptr: int, offset: int, length: int, seed: int, stream: int = 0):
    _check(lib().rrx_corpus_records_device(d_ptr, offset, length, seed, stream))


def corpus_lines(d_ptr: int, offset: int, length: int, seed: int, stream: int = 0):
    _check(lib().rrx_corpus_lines_device(d_ptr, offset, length, seed, stream))


def corpus_device(cfg, d_ptr: int, offset: int, length: int, stream: int = 0):
    """Bytes [offset, offset+length) of a configuration's corpus (configs.py), generated in device memory."""
    if cfg.mode == "iid":
        corpus_iid(d_ptr, offset, length, cfg.seed, cfg.alpha, stream)
    elif cfg.mode == "lines":
        corpus_lines(d_ptr, offset, length, cfg.seed, stream)
    else:
        corpus_records(d_ptr, offset, length, cfg.seed, stream)


class NcclComm:
    """A raw NCCL communicator (ncclComm_t) for rrx_sharded_*, made with libnccl.so.2 directly (ctypes): rank 0 draws
    the unique id, `broadcast(bytes) -> bytes` (any side channel: torch.distributed, MPI, a file) hands it to the others.
    The library is the one already in the process if there is one (e.g. torch's), which is also what librrx picks up."""

    def __init__(self, rank: int, world: int, broadcast):
        self.lib = C.CDLL("libnccl.so.2")
        uid = C.create_string_buffer(128)
        if rank == 0:
            assert self.lib.ncclGetUniqueId(uid) == 0
        raw = broadcast(bytes(uid.raw))
        uid = C.create_string_buffer(raw, 128)

        class _Uid(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]
        u = _Uid()
        C.memmove(C.byref(u), uid, 128)
        self.comm = C.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _Uid, C.c_int]
        rc = self.lib.ncclCommInitRank(C.byref(self.comm), world, u, rank)
        if rc != 0:
            raise RrxError(f"ncclCommInitRank failed ({rc})")
        self.rank, self.world = rank, world

    def close(self):
        if self.comm:
            self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


class ShardedScanner:
    """rrx_sharded_* (include/rrx.h): the sharded scan of one rank's shard, exchange over NCCL included."""

    def __init__(self, regex: RRegex, max_len: int, capacity: int, comm: NcclComm):
        self.regex, self.comm = regex, comm
        h = C.c_void_p()
        _check(lib().rrx_sharded_create(regex._h, max_len, capacity, comm.comm, comm.rank, comm.world, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _LIB is not None:
            _LIB.rrx_sharded_free(h)

    def begin(self, d_ptr: int, length: int, base_offset: int, left_context: int = 0, stream: int = 0):
        _check(lib().rrx_sharded_begin(self._h, d_ptr, length, base_offset, left_context, stream))

    def finish(self) -> dict:
        r = _ShardedResult()
        _check(lib().rrx_sharded_finish(self._h, C.byref(r)))
        return dict(n_matches=int(r.n_matches), d_matches=int(r.d_matches or 0), rescans=int(r.rescans),
                    patched_starts=int(r.patched_starts), flags=int(r.flags), scanner=r.scanner)

    def scan(self, d_ptr: int, length: int, base_offset: int, left_context: int = 0, stream: int = 0) -> dict:
        self.begin(d_ptr, length, base_offset, left_context, stream)
        return self.finish()

    def matches(self, res: dict, stream: int = 0) -> np.ndarray:
        out = np.empty((res["n_matches"], 2), dtype=np.int64)
        _check(lib().rrx_scanner_copy_matches(res["scanner"], out.ctypes.data, res["n_matches"], stream))
        return out

    def digest(self, res: dict, end_limit: int = (1 << 62), stream: int = 0) -> tuple:
        d = np.zeros(4, dtype=np.uint64)
        _check(lib().rrx_scanner_digest(res["scanner"], end_limit, d.ctypes.data, stream))
        return tuple(int(x) for x in d)

    def enable_timing(self, on: bool = True):
        _check(lib().rrx_sharded_enable_timing(self._h, int(on)))

    def kernel_ms(self) -> float:
        return float(lib().rrx_sharded_kernel_ms(self._h))
