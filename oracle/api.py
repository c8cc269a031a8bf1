"""TEST INFRASTRUCTURE ONLY — ctypes front end to oracle/_build/liboracle.so and oracle/_ref/."""
from __future__ import annotations

import ctypes as C
import functools
import os
import subprocess

import numpy as np

from . import ref_nfa

HERE = os.path.dirname(os.path.abspath(__file__))
LIBORACLE = os.path.join(HERE, "_build", "liboracle.so")
LIBREF = os.path.join(HERE, "_ref", "libref.so")
REFTOOL = os.path.join(HERE, "_ref", "ref_tool")


def build(ref: bool = True) -> None:
    """Build liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class _RawNFA(C.Structure):
    _fields_ = [("n", C.c_uint32), ("W", C.c_uint32), ("init", C.c_uint32),
                ("table", C.c_void_p), ("finals", C.c_void_p)]


@functools.lru_cache(maxsize=None)
def _lib():
    if not os.path.exists(LIBORACLE):
        build(ref=False)
    lib = C.CDLL(LIBORACLE)
    lib.oracle_accepts.restype = C.c_int
    lib.oracle_accepts.argtypes = [C.POINTER(_RawNFA), C.c_void_p, C.c_uint64]
    lib.oracle_lazy_scan.restype = C.c_uint64
    lib.oracle_lazy_scan.argtypes = [C.POINTER(_RawNFA), C.c_void_p, C.c_uint64, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.oracle_time_parallel.restype = C.c_double
    lib.oracle_time_parallel.argtypes = [C.POINTER(_RawNFA), C.c_void_p, C.c_uint64, C.c_int,
                                         C.POINTER(C.c_uint64)]
    lib.corpus_fill_iid.restype = None
    lib.corpus_fill_iid.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.corpus_fill_records.restype = None
    lib.corpus_fill_records.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.corpus_fill_lines.restype = None
    lib.corpus_fill_lines.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.corpus_lines_valid.restype = C.c_uint64
    lib.corpus_lines_valid.argtypes = [C.c_uint64, C.c_uint64]
    lib.corpus_records_valid.restype = C.c_uint64
    lib.corpus_records_valid.argtypes = [C.c_uint64, C.c_uint64]
    lib.corpus_h.restype = C.c_uint64
    lib.corpus_h.argtypes = [C.c_uint64, C.c_uint64]
    return lib


@functools.lru_cache(maxsize=None)
def _ref():
    if not os.path.exists(LIBREF):
        return None
    lib = C.CDLL(LIBREF)
    lib.ref_accepts.restype = C.c_int
    lib.ref_accepts.argtypes = [C.c_char_p, C.c_char_p]
    lib.ref_time_acceptance.restype = C.c_double
    lib.ref_time_acceptance.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    return lib


def have_ref() -> bool:
    return _ref() is not None


class Oracle:
    """A compiled pattern: raw reference-layout tables + the C scanner."""

    def __init__(self, pattern):
        self.pattern = pattern
        self.raw = ref_nfa.compile_raw(pattern)
        self._table, self._finals = self.raw.export_table()
        self._s = _RawNFA(self.raw.n, self.raw.words(), self.raw.init,
                          self._table.ctypes.data, self._finals.ctypes.data)

    @property
    def n_raw(self) -> int:
        return self.raw.n

    def accepts(self, text: bytes) -> bool:
        buf = np.frombuffer(bytes(text), dtype=np.uint8) if len(text) else np.zeros(0, np.uint8)
        return bool(_lib().oracle_accepts(C.byref(self._s), buf.ctypes.data, len(text)))

    def lazy_matches(self, text, base: int = 0) -> np.ndarray:
        """LAZY-v1 match list as an (M,2) int64 array of half-open (start,end) offsets."""
        buf = _as_u8(text)
        n = len(buf)
        cap = max(1024, n // 64)
        while True:
            out = np.empty((cap, 2), dtype=np.int64)
            m = _lib().oracle_lazy_scan(C.byref(self._s), buf.ctypes.data, n, base, None,
                                        out.ctypes.data, cap, None)
            if m <= cap:
                return out[:m].copy()
            cap = int(m)

    def lazy_matches_anchored(self, text) -> np.ndarray:
        """LAZY-v1 with working anchors (the extension DESIGN.md §8 defines; the reference never matches them,
        regex.h:157): the text is framed by ONE virtual 0x00 on each side which only the anchor atoms consume; the list
        of the framed text is mapped back to text coordinates (virtual bytes clipped), empty matches dropped, one match
        per end (the one with the larger start).  The text must not contain 0x00."""
        buf = _as_u8(text)
        n = len(buf)
        assert not (buf == 0).any()
        framed = np.concatenate([np.zeros(1, np.uint8), buf, np.zeros(1, np.uint8)])
        o = Oracle.__new__(Oracle)
        o.pattern, o.raw = self.pattern, self.raw.anchored_view()
        o._table, o._finals = o.raw.export_table()
        o._s = _RawNFA(o.raw.n, o.raw.words(), o.raw.init, o._table.ctypes.data, o._finals.ctypes.data)
        m = o.lazy_matches(framed, base=-1)
        if len(m) == 0:
            return m
        s, e = np.maximum(m[:, 0], 0), np.minimum(m[:, 1], n)
        best = {}
        for a, b in zip(s.tolist(), e.tolist()):
            if a < b and best.get(b, -1) < a:
                best[b] = a
        return np.array([[best[b], b] for b in sorted(best)], dtype=np.int64).reshape(-1, 2)

    def no_overlap_matches(self, text, base: int = 0) -> np.ndarray:
        """The non-overlapping view of the LAZY-v1 list."""
        m = np.ascontiguousarray(self.lazy_matches(text, base))
        L = _lib()
        L.oracle_no_overlap.restype = C.c_uint64
        L.oracle_no_overlap.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        k = int(L.oracle_no_overlap(m.ctypes.data, len(m), base))
        return m[:k].copy()

    def time_parallel(self, text, piece: int, nthreads: int):
        buf = _as_u8(text)
        assert piece * nthreads <= len(buf)
        tot = C.c_uint64(0)
        sec = _lib().oracle_time_parallel(C.byref(self._s), buf.ctypes.data, piece, nthreads, C.byref(tot))
        return sec, int(tot.value)


def _as_u8(text) -> np.ndarray:
    if isinstance(text, np.ndarray):
        assert text.dtype == np.uint8
        return np.ascontiguousarray(text)
    if isinstance(text, str):
        text = text.encode("latin-1")
    return np.frombuffer(bytes(text), dtype=np.uint8) if len(text) else np.zeros(0, np.uint8)


def ref_accepts(pattern: str, text: bytes) -> bool:
    """Whole-string acceptance by the compiled, unmodified reference (needs oracle/_ref)."""
    lib = _ref()
    assert lib is not None, "oracle/_ref/libref.so not built"
    assert b"\0" not in text
    r = lib.ref_accepts(pattern.encode("latin-1"), bytes(text))
    if r < 0:
        raise ref_nfa.InvalidExpression(pattern)
    return bool(r)


def ref_time_acceptance(pattern: str, text: bytes, repeats: int = 3):
    lib = _ref()
    assert lib is not None
    acc = C.c_int(0)
    sec = lib.ref_time_acceptance(pattern.encode("latin-1"), bytes(text), repeats, C.byref(acc))
    return sec, bool(acc.value)


def ref_dump(pattern: str):
    """Parse the reference's own NFA::print() output (NFA.cc:14-41) into (n, finals, fwd, bwd)
    with fwd/bwd = {state: {byte: [targets]}}."""
    import re
    out = subprocess.run([REFTOOL, "dump", pattern], check=True, capture_output=True).stdout

    def parse_set(inner: bytes):
        return sorted(int(x) for x in inner.replace(b"\x08", b"").split(b",") if x.strip())

    head = re.match(rb"initial state: (\d+)\nbuffer_size: (\d+)\nsize: (\d+)\nfinal states: \{([0-9,\x08]*)\}\n", out)
    n = int(head.group(3))
    finals = parse_set(head.group(4))
    fwd, bwd = {}, {}
    tok = re.compile(rb"state: (\d+)\n|(forward|backward) transitions: \n|([\x00-\x7f]): \{([0-9,\x08]*)\}\n", re.S)
    pos, state, cur = head.end(), None, None
    while pos < len(out):
        m = tok.match(out, pos)
        assert m, out[pos:pos + 40]
        if m.group(1) is not None:
            state = int(m.group(1))
        elif m.group(2) is not None:
            cur = fwd if m.group(2) == b"forward" else bwd
        else:
            cur.setdefault(state, {})[m.group(3)[0]] = parse_set(m.group(4))
        pos = m.end()
    return n, finals, fwd, bwd


def _sm64(z: np.ndarray) -> np.ndarray:
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def digest(matches: np.ndarray, end_limit: int = (1 << 62)) -> tuple:
    """(count, sum, xor, order_violations) of a host match list — the CPU twin of csrc/rrx_hash.cu
    (count + sum/xor of splitmix over the (start,end) pairs)."""
    m = np.ascontiguousarray(matches, dtype=np.int64).reshape(-1, 2)
    bad = int(np.count_nonzero(m[:, 0] >= m[:, 1])) + (int(np.count_nonzero(m[1:, 1] <= m[:-1, 1])) if len(m) > 1 else 0)
    m = m[m[:, 1] <= end_limit]
    if len(m) == 0:
        return (0, 0, 0, bad)
    with np.errstate(over="ignore"):
        h = _sm64(_sm64(m[:, 0].astype(np.uint64)) + m[:, 1].astype(np.uint64))
        s = int(np.add.reduce(h, dtype=np.uint64))
    return (len(m), s, int(np.bitwise_xor.reduce(h)), bad)


def corpus_iid(seed: int, alpha: bytes, offset: int, length: int) -> np.ndarray:
    assert len(alpha) == 256
    out = np.empty(length, dtype=np.uint8)
    a = np.frombuffer(bytes(alpha), dtype=np.uint8)
    _lib().corpus_fill_iid(seed, a.ctypes.data, offset, length, out.ctypes.data)
    return out


def corpus_records(seed: int, offset: int, length: int) -> np.ndarray:
    out = np.empty(length, dtype=np.uint8)
    _lib().corpus_fill_records(seed, offset, length, out.ctypes.data)
    return out


def corpus_lines(seed: int, offset: int, length: int) -> np.ndarray:
    """Config 2 variant: unpadded variable-length log lines (oracle/corpus.c corpus_fill_lines)."""
    out = np.empty(length, dtype=np.uint8)
    _lib().corpus_fill_lines(seed, offset, length, out.ctypes.data)
    return out


def corpus(cfg, offset: int, length: int) -> np.ndarray:
    """Bytes [offset, offset+length) of a configuration's corpus (roaringregex_b200/configs.py)."""
    if cfg.mode == "iid":
        return corpus_iid(cfg.seed, cfg.alpha, offset, length)
    if cfg.mode == "lines":
        return corpus_lines(cfg.seed, offset, length)
    return corpus_records(cfg.seed, offset, length)


def lines_valid(seed: int, n_blocks: int) -> int:
    return int(_lib().corpus_lines_valid(seed, n_blocks))


def records_valid(seed: int, n_records: int) -> int:
    return int(_lib().corpus_records_valid(seed, n_records))
