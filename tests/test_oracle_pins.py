"""Pins the CPU oracle (oracle/) to the reference: golden vectors generated from the compiled,
unmodified reference (tests/golden/make_golden.py), the live compiled reference when oracle/_ref is
present, and Python `re` for quirk-free patterns.  CPU only."""
import json
import os
import random
import re

import numpy as np
import pytest

from oracle import ref_nfa

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DUMPS = json.load(open(os.path.join(GOLD, "nfa_dumps.json")))
ACCEPT = json.load(open(os.path.join(GOLD, "acceptance.json")))
LAZY = json.load(open(os.path.join(GOLD, "lazy_v1.json")))


def _norm(rows):
    return {int(s): {int(c): t for c, t in r.items()} for s, r in rows.items()}


@pytest.mark.parametrize("pattern", sorted(DUMPS))
def test_construction_reproduces_reference_tables(pattern):
    """Appendix B.1: the restated construction yields the reference's own printed tables
    (raw state count, final states, forward AND backward rows)."""
    g = DUMPS[pattern]
    r = ref_nfa.compile_raw(pattern)
    assert r.n == g["n"]
    assert sorted(ref_nfa._bits(r.finals)) == g["finals"]
    assert r.dump_forward() == _norm(g["fwd"])
    assert r.dump_backward() == _norm(g["bwd"])


def test_raw_state_counts_of_the_configs():
    # SURVEY.md §0.5 [probe]: 10 / 30 / 301 / 1501 / 10
    for pat, n in [("(ab|cd)*e", 10), ("[a-z]+@[a-z]+\\.(com|org|net)", 30), ("a{1,100}b", 301),
                   ("(a|b){1,300}c", 1501), ("(x|y)*z.*q", 10)]:
        assert ref_nfa.compile_raw(pat).n == n


def test_acceptance_golden(oracle_api):
    """Appendix B.2 + 780 seeded random strings answered by the reference's AcceptanceIterator."""
    cache = {}
    for pat, text, want in ACCEPT:
        o = cache.get(pat) or cache.setdefault(pat, oracle_api.Oracle(pat))
        t = text.encode("latin-1")
        assert o.accepts(t) == bool(want), (pat, text)
        assert ref_nfa.accepts(o.raw, t) == bool(want), (pat, text)


def test_lazy_v1_golden(oracle_api):
    """Appendix B.3 + random: brute force over (start,end) with the reference as membership test."""
    cache = {}
    for pat, text, want in LAZY:
        o = cache.get(pat) or cache.setdefault(pat, oracle_api.Oracle(pat))
        got = o.lazy_matches(text.encode("latin-1")).tolist()
        assert got == want, (pat, text)
        assert [list(x) for x in ref_nfa.lazy_matches_py(o.raw, text.encode("latin-1"))] == want


def test_lazy_v1_survey_vectors_large_automata(oracle_api):
    # Appendix B.3 rows whose raw state count (> 256) the reference itself cannot run
    assert oracle_api.Oracle("a{1,100}b").lazy_matches(b"aaab b ab aab").tolist() == [[2, 4], [7, 9], [11, 13]]
    assert oracle_api.Oracle("(a|b){1,300}c").lazy_matches(b"abbac c bc").tolist() == [[3, 5], [8, 10]]
    o = oracle_api.Oracle("a{1,100}b")
    assert o.accepts(b"a" * 100 + b"b") and not o.accepts(b"a" * 101 + b"b") and not o.accepts(b"b")
    o = oracle_api.Oracle("(a|b){1,300}c")
    assert not o.accepts(b"c") and o.accepts(b"ab" * 150 + b"c") and not o.accepts(b"ab" * 150 + b"ac")


def test_live_reference_acceptance(oracle_api):
    """When the compiled reference is present: fresh random strings, not just the golden ones."""
    if not oracle_api.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    rnd = random.Random(7)
    for pat, alpha in [("(ab|cd)*e", "abcde"), ("(x|y)*z.*q", "xyzq\n"), ("a{1,80}b", "ab"), ("(a|b){1,40}c", "abc"),
                       ("[a-z]+@[a-z]+\\.(com|org|net)", "comnetrg@."), ("x*|y", "xy"), ("[^a]+b", "ab^")]:
        o = oracle_api.Oracle(pat)
        for _ in range(150):
            t = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 90))).encode()
            assert o.accepts(t) == oracle_api.ref_accepts(pat, t), (pat, t)


def test_lazy_against_python_re(oracle_api):
    """Quirk-free patterns: ENDS and the shortest match per end from Python's re (fullmatch, DOTALL)."""
    rnd = random.Random(11)
    for pat, alpha in [("(ab|cd)*e", "abcde "), ("[a-z]+@[a-z]+\\.(com|org|net)", "acomnetrg@. "), ("a{1,100}b", "aaab "),
                       ("(a|b){1,300}c", "abc "), ("(x|y)*z.*q", "xyzq a\n")]:
        o = oracle_api.Oracle(pat)
        rx = re.compile(pat, re.S)
        for _ in range(20):
            t = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 40)))
            want = []
            for e in range(1, len(t) + 1):
                for s in range(e - 1, -1, -1):
                    if rx.fullmatch(t, s, e):
                        want.append([s, e])
                        break
            assert o.lazy_matches(t.encode()).tolist() == want, (pat, t)


def test_anchored_oracle_against_python_re(oracle_api):
    """The anchored extension of the oracle (virtual 0x00 bytes that only anchor atoms consume) against a brute force
    with Python `re`: `^X` = X matching from offset 0 only, `X$` = X matching up to the text's end only."""
    import random
    import re
    rnd = random.Random(3)
    for body, alpha in [("ab", b"ab "), ("a+b", b"ab "), ("(ab|cd)*e", b"abcde "), ("[a-c]+x", b"abcx "), ("a.c", b"abc\n")]:
        rx = re.compile(body.encode(), re.S)
        for _ in range(40):
            n = rnd.randint(0, 40)
            t = bytes(rnd.choices(alpha, k=n))
            for kind, pat in (("^", "^" + body), ("$", body + "$"), ("^$", "^" + body + "$")):
                want = []
                for e in range(1, n + 1):
                    if "$" in kind and e != n:
                        continue
                    starts = [0] if "^" in kind else range(e - 1, -1, -1)
                    for s in starts:
                        if s < e and rx.fullmatch(t[s:e]):
                            want.append([s, e])
                            break
                got = oracle_api.Oracle(pat).lazy_matches_anchored(t).tolist()
                assert got == want, (pat, t, got, want)
