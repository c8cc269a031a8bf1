"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ERE -> NFA construction.

This module restates, on RAW reference states, what `/root/reference/src/Parser.cpp:16-170` and
`/root/reference/src/NFA.cc:42-71,108-189` build: the mirrored forward/backward transition tables
`table[byte][direction][state] -> set of states` (layout of `NFA.cc:9-12`, *without* the
`(uint8_t)` truncation of the state index), the initial state and the set
of final states.  It deliberately works on the reference's raw state numbering (two states per
atom, one per `?`), NOT on the flattened position automaton the CUDA path uses, so that the
product's host compiler (roaringregex_b200/csrc/rrx_compile.cc, which tracks atoms) and this
oracle are two independent derivations of the same language.

Pinned by: tests/test_oracle_pins.py — the NFA dumps printed by the reference's own
`NFA::print()` (regenerated from oracle/_ref/ref_tool), whole-string
acceptance of the compiled reference (oracle/_ref/libref.so) on random strings, and Python `re`
for quirk-free patterns.

State sets are Python ints used as bitsets (bit i = raw state i).
Only tests/, bench.py's cpu_baseline/reference arm and __graft_entry__.smoke() may import this.
"""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field

import numpy as np

CONCATENATION, BRACKETS, OR = 0, 1, 2  # regex.h:15
ALL128 = (1 << 128) - 1


class InvalidExpression(ValueError):
    """Mirrors the two `throw std::runtime_error` sites, Parser.cpp:36 and :155."""


def _bits(x: int):
    while x:
        low = x & -x
        yield low.bit_length() - 1
        x ^= low


@dataclass
class _Sub:
    """A sub-automaton on the parser's stack: PseudoNFA{initial_state,size} + final_states."""
    init: int
    size: int
    finals: int


@dataclass
class RawNFA:
    n: int                      # number of raw states (states_n, Parser.cpp:163)
    init: int                   # always 0
    finals: int                 # bitset
    fwd: dict = field(repr=False, default_factory=dict)   # (state, byte) -> bitset of targets
    bwd: dict = field(repr=False, default_factory=dict)   # (state, byte) -> bitset of sources

    @property
    def width_class(self) -> str:
        """Set type the reference would pick, Parser.cpp:165-168."""
        if self.n > 256:
            return "Roaring"
        if self.n > 128:
            return "BitSet<4>"
        if self.n > 64:
            return "BitSet<2>"
        return "BitSet<1>"

    def words(self) -> int:
        return (self.n + 63) // 64

    def export_table(self):
        """Dense table in the reference's layout (NFA.cc:9-12): row = state + n*(2*byte + dir),
        dir 0 = forward, 1 = backward; each row is `words()` uint64.  Returns (table, finals)."""
        W = self.words()
        tab = np.zeros((128 * 2 * self.n, W), dtype=np.uint64)
        mask = (1 << 64) - 1
        for d, src in ((0, self.fwd), (1, self.bwd)):
            for (s, c), v in src.items():
                if not v:
                    continue
                row = s + self.n * (2 * c + d)
                for w in range(W):
                    tab[row, w] = (v >> (64 * w)) & mask
        fin = np.array([(self.finals >> (64 * w)) & mask for w in range(W)], dtype=np.uint64)
        return tab, fin

    def anchored_view(self) -> "RawNFA":
        """The same automaton with its byte-0x00 transitions restricted to ANCHOR atoms (`^`, `$`: Parser.cpp:142-146
        compiles both to a literal 0x00), for scanning a text framed by virtual 0x00 bytes: `.` and negated classes,
        which also contain 0x00, must not consume those.  An anchor atom's own final state t* is entered on byte 0 and
        on no other byte; the construction (skip<bwd>, NFA.cc:108-137) copies every transition into t* to the states
        that follow the atom, so a byte-0 edge s -> T belongs to the anchor iff s enters t* on byte 0 and every source
        of t* on byte 0 also enters T on byte 0.  (Limitation: a broad class in alternation with an anchor AND with the
        same continuation — `(^|.)a` — cannot be told apart from the tables alone and keeps its edge; the tests avoid it.)"""
        incoming = defaultdict(int)
        for (s, c), v in self.fwd.items():
            for t in _bits(v):
                incoming[t] |= 1 << c
        own = [t for t, m in incoming.items() if m == 1]
        pre = {t: self.bwd.get((t, 0), 0) for t in own}
        fwd, bwd = {}, defaultdict(int)
        for (s, c), v in self.fwd.items():
            if c == 0:
                keep = 0
                for T in _bits(v):
                    srcT = self.bwd.get((T, 0), 0)
                    if any((p >> s) & 1 and (p & ~srcT) == 0 for p in pre.values()):
                        keep |= 1 << T
                v = keep
            if v:
                fwd[(s, c)] = v
                for t in _bits(v):
                    bwd[(t, c)] |= 1 << s
        return RawNFA(self.n, self.init, self.finals, fwd, dict(bwd))

    def dump_forward(self):
        """{state: {byte: sorted targets}} for non-empty forward rows (what NFA::print shows)."""
        out = defaultdict(dict)
        for (s, c), v in self.fwd.items():
            if v:
                out[s][c] = sorted(_bits(v))
        return dict(out)

    def dump_backward(self):
        out = defaultdict(dict)
        for (s, c), v in self.bwd.items():
            if v:
                out[s][c] = sorted(_bits(v))
        return dict(out)


class _Builder:
    """The construction algebra of NFA.cc on one shared pair of tables (the MemoryPool)."""

    def __init__(self):
        self.fwd = defaultdict(int)
        self.bwd = defaultdict(int)

    # --- the three constructors, NFA.cc:42-71 ------------------------------------------------
    def empty(self, cur):
        return _Sub(cur, 1, 1 << cur)

    def atom(self, cur, charset):
        """`charset` is a 128-bit set of byte values; a literal is a one-element set."""
        for c in _bits(charset):
            self.fwd[(cur, c)] |= 1 << (cur + 1)
            self.bwd[(cur + 1, c)] |= 1 << cur
        return _Sub(cur, 2, 1 << (cur + 1))

    # --- skip<fwd>(n,k), NFA.cc:108-121 ---------------------------------------------------------
    def skip(self, forward, n, k):
        to_skip, fixed = (k, n) if forward else (n, k)
        tab, other = (self.fwd, self.bwd) if forward else (self.bwd, self.fwd)
        for c in range(128):
            row = tab.get((to_skip, c), 0)
            if not row:
                continue
            tab[(fixed, c)] |= row
            for j in _bits(row):
                other[(j, c)] |= 1 << fixed

    # --- concatenation, NFA.cc:122-137 ------------------------------------------------------------
    def concat(self, a, b):
        a.size += b.size
        for f in list(_bits(a.finals)):
            self.skip(False, f, b.init)
        a_nullable = bool(a.finals >> a.init & 1)
        if a_nullable:
            self.skip(True, a.init, b.init)
        if a_nullable and (b.finals >> b.init & 1):
            a.finals = b.finals | (1 << a.init)
        else:
            a.finals = b.finals
        return a

    # --- union, NFA.cc:138-149 ----------------------------------------------------------------------
    def union(self, a, b):
        a.size += b.size
        a.finals |= b.finals
        self.skip(True, a.init, b.init)
        if b.finals >> b.init & 1:
            a.finals |= 1 << a.init
        return a

    # --- Kleene star, NFA.cc:150-157 ----------------------------------------------------------------
    def star(self, a):
        for f in list(_bits(a.finals)):
            self.skip(False, f, a.init)
        a.finals |= 1 << a.init
        return a

    # --- relocated copy, NFA.cc:177-185 (+ set shift, BitSet.cc:168-180 / NFA.cc:172-176) -----------
    def clone(self, a, r):
        for i in range(a.init, a.init + a.size):
            for c in range(128):
                for tab in (self.fwd, self.bwd):
                    v = tab.get((i, c), 0)
                    if v or (i + r, c) in tab:
                        tab[(i + r, c)] = v << r
        return _Sub(a.init + r, a.size, a.finals << r)


def _strtol(p: bytes, pos: int):
    """C strtol(…, 10) on a NUL-terminated buffer: returns (value, endpos)."""
    i = pos
    n = len(p)
    while i < n and p[i] in b" \t\n\v\f\r":
        i += 1
    neg = False
    if i < n and p[i] in b"+-":
        neg = p[i] == ord("-")
        i += 1
    start = i
    val = 0
    while i < n and 48 <= p[i] <= 57:
        val = val * 10 + (p[i] - 48)
        i += 1
    if i == start:
        return 0, pos
    return (-val if neg else val), i


def _bracket(p: bytes, pos: int):
    """Bracket expression, Parser.cpp:16-39.  `pos` is at '['.  Returns (charset, pos_of_end)."""
    end = len(p)

    def at(i):
        return p[i] if i < end else 0

    start = pos
    ps = end - start
    pos += 1
    complement = at(pos) == ord("^")     # the caret is NOT skipped: it lands in the set (quirk A.3.3)
    charset = 0
    escaped = False
    while pos + 1 < end and not (at(pos) == ord("]") and not escaped):
        if not escaped:
            nxt = at(pos + 1)
            if nxt != ord("]") and pos + 2 < end:
                nn = at(pos + 2)
                if nxt == ord("-") and nn != ord("]"):
                    for c in range(at(pos), nn + 1):
                        charset |= 1 << c
                    pos += 3
                    continue
        escaped = (not escaped) and at(pos) == ord("\\")
        charset |= 1 << at(pos)          # a backslash is itself added (quirk A.3.4)
        pos += 1
    if pos - start == ps:
        raise InvalidExpression("invalid expression!")
    if complement:
        charset ^= ALL128
    return charset, pos


def compile_raw(pattern) -> RawNFA:
    """RRegex::RRegex + build_NFA, Parser.cpp:41-170."""
    p = pattern.encode("latin-1") if isinstance(pattern, str) else bytes(pattern)
    if any(b >= 0x80 or b == 0 for b in p):
        raise InvalidExpression("pattern must be 7-bit ASCII without NUL (reference: UB)")
    ps = len(p)
    b = _Builder()
    nfas: list[_Sub] = []
    ops: list[int] = [BRACKETS]

    def next_initial():
        return 0 if not nfas else nfas[-1].init + nfas[-1].size

    def need_top():
        if not nfas:
            raise InvalidExpression("operator without operand (reference: UB)")

    def clear_stack():
        # Parser.cpp:49-79 — fold the stack right-to-left down to the innermost open bracket.
        need_top()
        cur = nfas[-1]
        if ops:
            ops.pop()
            while len(ops) > 1 and ops[-1] != BRACKETS:
                if ops[-1] == CONCATENATION:
                    nfas.pop()
                    ops.pop()
                    need_top()
                    cur = b.concat(nfas[-1], cur)
                elif ops[-1] == OR:
                    inter = cur
                    ops.pop()
                    if not ops:
                        raise InvalidExpression("invalid expression")
                    ops.pop()
                    nfas.pop()
                    need_top()
                    cur = nfas[-1]
                    while len(ops) > 1 and ops[-1] == CONCATENATION:
                        nfas.pop()
                        ops.pop()
                        need_top()
                        cur = b.concat(nfas[-1], cur)
                    cur = b.union(cur, inter)
        if not ops:
            raise InvalidExpression("invalid expression")
        ops.pop()
        ops.append(CONCATENATION)
        nfas.pop()
        nfas.append(cur)

    def repeat():
        need_top()
        top = nfas[-1]
        nfas.append(b.clone(top, top.size))
        ops.append(CONCATENATION)

    def at(i):
        return p[i] if i < ps else 0

    cp = 0
    escaped = False
    while True:                       # do { … } while (++cp - p < ps)
        ch = at(cp)
        if not escaped and ch == ord("\\"):
            escaped = True
            cp += 1
            if cp < ps:
                continue
            break
        sel = 0 if escaped else ch
        if sel == ord("["):
            cs, cp = _bracket(p, cp)
            nfas.append(b.atom(next_initial(), cs))
            ops.append(CONCATENATION)
        elif sel == ord("("):
            ops.append(BRACKETS)
        elif sel == ord("|"):
            ops.append(OR)
        elif sel == ord(")"):
            clear_stack()
        elif sel == ord("."):
            nfas.append(b.atom(next_initial(), ALL128))
            ops.append(CONCATENATION)
        elif sel == ord("*"):
            need_top()
            b.star(nfas[-1])
        elif sel == ord("+"):
            repeat()
            b.star(nfas[-1])
        elif sel == ord("?"):
            need_top()
            b.union(nfas[-1], b.empty(next_initial()))
        elif sel == ord("{"):
            cp += 1
            m, cn1 = _strtol(p, cp)
            for _ in range(m - 1):
                repeat()
            if at(cn1) != ord("}"):
                n, cn2 = _strtol(p, cn1 + 1)
                if not n:
                    repeat()
                    b.star(nfas[-1])
                elif n > m:
                    repeat()
                    b.union(nfas[-1], b.empty(next_initial()))
                    for _ in range(m + 1, n):
                        repeat()
                cp = cn2
            else:
                cp = cn1
        elif sel in (ord("^"), ord("$")):
            nfas.append(b.atom(next_initial(), 1 << 0))   # both anchors: literal byte 0x00
            ops.append(CONCATENATION)
        else:
            nfas.append(b.atom(next_initial(), 1 << ch))
            ops.append(CONCATENATION)
        escaped = False
        cp += 1
        if not cp < ps:
            break
    clear_stack()
    if len(nfas) != 1:
        raise InvalidExpression("invalid expression")
    top = nfas[-1]
    return RawNFA(n=top.size, init=top.init, finals=top.finals,
                  fwd={k: v for k, v in b.fwd.items() if v},
                  bwd={k: v for k, v in b.bwd.items() if v})


def accepts(nfa: RawNFA, text: bytes) -> bool:
    """AcceptanceIterator (regex.h:156-162) with Processor::shift<true> (NFA.cc:72-102):
    anchored whole-string membership, no re-injection.  Pure Python; small inputs only."""
    cur = 1 << nfa.init
    for c in text:
        if c >= 128:
            return False
        nxt = 0
        for i in _bits(cur):
            nxt |= nfa.fwd.get((i, c), 0)
        cur = nxt
        if not cur:
            return False
    return bool(cur & nfa.finals)


def lazy_matches_py(nfa: RawNFA, text: bytes):
    """LAZY-v1 in pure Python; small inputs only.  The fast version of
    the same loop is oracle/lazy_oracle.c."""
    out = []
    S = 0
    for i, c in enumerate(text):
        S |= 1 << nfa.init
        nxt = 0
        if c < 128:
            for s in _bits(S):
                nxt |= nfa.fwd.get((s, c), 0)
        S = nxt
        if S & nfa.finals:
            Bk = nfa.finals
            j = i + 1
            while True:
                j -= 1
                cj = text[j]
                nb = 0
                if cj < 128:
                    for s in _bits(Bk):
                        nb |= nfa.bwd.get((s, cj), 0)
                Bk = nb
                if Bk >> nfa.init & 1:
                    break
                assert Bk and j > 0, "forward accept without a backward path"
            out.append((j, i + 1))
    return out
