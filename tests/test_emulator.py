"""The kernel's tile logic (roaringregex_b200/csrc/rrx_tile.cuh) executed on the CPU by the test
emulator (tests/emu/rrx_emu.cc) against the oracle: row transfer functions, look-back, prefix replay,
lazy start scans, reductions of the device automaton.  CPU only; the same functions run on the GPU."""
import random

import numpy as np
import pytest

import emu_api
from conftest import cpu_corpus
from roaringregex_b200.configs import CONFIGS


def _check(oracle_api, pat, text, **kw):
    want = oracle_api.Oracle(pat).lazy_matches(text) if not isinstance(pat, tuple) else pat[1].lazy_matches(text)
    p = pat if not isinstance(pat, tuple) else pat[0]
    for reduce, fast, donly in ((True, True, True), (True, True, False), (False, False, True), (True, False, False), (False, True, False)):
        # donly: let the compiler pick the thread-owned candidate scan where it applies / forbid it
        r = emu_api.scan(p, text, reduce=reduce, fast=fast, donly=donly, **kw)
        assert r["err"] == 0
        assert r["n"] == len(want), (p, reduce, fast, donly, r["n"], len(want))
        assert np.array_equal(r["matches"], want), (p, reduce, fast, donly)


@pytest.mark.parametrize("cfg_id", [1, 2, 3, 4, 5])
def test_config_corpora_small(oracle_api, cfg_id):
    cfg = CONFIGS[cfg_id]
    n = 200_000 if cfg_id != 4 else 100_000
    text = cpu_corpus(cfg, 0, n)
    _check(oracle_api, cfg.pattern, text)


def test_survey_vectors(oracle_api):
    for pat, txt in [("(ab|cd)*e", "xxababe cde e"), ("[a-z]+@[a-z]+\\.(com|org|net)", "foo@bar.com x a@b.org c@d.comet@e.net"),
                     ("a{1,100}b", "aaab b ab aab"), ("(a|b){1,300}c", "abbac c bc"), ("(x|y)*z.*q", "xyzaaqbbq q zq"),
                     ("ab|bc", "abc"), ("a+", "aaa")]:
        _check(oracle_api, pat, txt.encode())


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 127, 128, 129, 255, 256, 4095, 4096, 4097, 32767, 32768, 32769, 65536 + 77])
def test_sizes_around_row_and_tile_edges(oracle_api, n):
    rnd = random.Random(n)
    for pat, alpha in [("(ab|cd)*e", b"abcde "), ("(x|y)*z.*q", b"xyzq a"), ("a{1,100}b", b"aaaaab ")]:
        text = bytes(rnd.choice(alpha) for _ in range(n))
        _check(oracle_api, pat, text)


def test_match_straddling_every_boundary(oracle_api):
    # one e-mail address slid across a row boundary (128) and a tile boundary (32768), every offset
    pat = "[a-z]+@[a-z]+\\.(com|org|net)"
    o = oracle_api.Oracle(pat)
    addr = b"someone@example.org"
    for edge in (128, 4096, 32768):
        for shift in range(len(addr) + 2):
            start = edge - shift
            text = bytearray(b" " * (edge + 64))
            text[start:start + len(addr)] = addr
            _check(oracle_api, (pat, o), bytes(text))


def test_unbounded_carry(oracle_api):
    # a single z at offset 0, the only q far away (beyond several tiles): the ".*" state must be
    # carried through every chunk; start scan walks back to offset 0
    n = 3 * 32768 + 500
    text = bytearray(b"a" * n)
    text[0] = ord("z")
    text[-1] = ord("q")
    text[70000] = ord("\n")           # '.' matches '\n' in this dialect
    _check(oracle_api, "(x|y)*z.*q", bytes(text))
    # no z at all: zero matches
    text[0] = ord("a")
    _check(oracle_api, "(x|y)*z.*q", bytes(text))


def test_hard_rows(oracle_api):
    # "(a|b){1,300}c": a run of letters longer than a row keeps > 1 survivor alive through whole
    # rows (transfer function of high rank): exercised with runs that cross rows and a tile edge
    rnd = random.Random(5)
    for run in (100, 128, 200, 299, 300, 301, 420, 1000):
        text = bytearray()
        while len(text) < 40000:
            text += bytes(rnd.choice(b"ab") for _ in range(run)) + b"c" + bytes(rnd.choice(b"ab cd") for _ in range(rnd.randint(0, 40)))
        _check(oracle_api, "(a|b){1,300}c", bytes(text))
    # all letters, one c at the very end, and nothing else
    _check(oracle_api, "(a|b){1,300}c", b"ab" * 20000 + b"c")
    # a pattern whose survivors never collapse and never die: carried as exact states tile to tile
    _check(oracle_api, "(a|b)*a(a|b){3}c", bytes(rnd.choice(b"ab") for _ in range(70000)) + b"c")


def _random_regex(rnd, depth=0):
    atoms = ["a", "b", "c", "[ab]", "[^a]", ".", "ab", "(a|b)", "(ab|c)"]
    if depth > 2:
        return rnd.choice(atoms)
    kind = rnd.randint(0, 7)
    if kind <= 1:
        return rnd.choice(atoms)
    if kind == 2:
        return _random_regex(rnd, depth + 1) + _random_regex(rnd, depth + 1)
    if kind == 3:
        return "(" + _random_regex(rnd, depth + 1) + "|" + _random_regex(rnd, depth + 1) + ")"
    if kind == 4:
        return "(" + _random_regex(rnd, depth + 1) + ")" + rnd.choice(["*", "+", "?"])
    if kind == 5:
        m = rnd.randint(1, 3)
        return "(" + _random_regex(rnd, depth + 1) + "){" + str(m) + rnd.choice(["", ",", "," + str(m + rnd.randint(1, 4))]) + "}"
    return _random_regex(rnd, depth + 1) + rnd.choice(atoms)


def test_random_regexes(oracle_api):
    rnd = random.Random(2026)
    done = 0
    while done < 120:
        pat = _random_regex(rnd)
        try:
            o = oracle_api.Oracle(pat)
        except Exception:
            continue
        if o.n_raw > 400:
            continue
        done += 1
        for n in (rnd.randint(0, 300), rnd.randint(300, 3000)):
            text = bytes(rnd.choice(b"abc\n") for _ in range(n))
            _check(oracle_api, (pat, o), text)


def test_sharded_scan_with_entry_and_left_context(oracle_api):
    """A shard scanned with the exact entry set and with left context gives exactly the matches of
    the whole text that end inside the shard; the shard summary reproduces the exit state."""
    rnd = random.Random(9)
    for pat, alpha in [("(x|y)*z.*q", b"xyzq a"), ("[a-z]+@[a-z]+\\.(com|org|net)", b"acomnetrg@. "), ("a{1,100}b", b"aaaab ")]:
        o = oracle_api.Oracle(pat)
        text = bytes(rnd.choice(alpha) for _ in range(4 * 32768))
        want = o.lazy_matches(text)
        for cut in (32768, 65536 + 128, 3 * 32768 - 128):
            left = emu_api.scan(pat, text[:cut])
            assert left["shard_comp"][0] & 0xff in (0, 1), "closed or rank-1 summary expected"
            NW = left["info"]["NW"]
            comp = left["shard_comp"]
            kind, b, U, Pm = comp[0] & 0xff, comp[1:1 + NW], comp[1 + NW:1 + 2 * NW], comp[1 + 2 * NW:1 + 3 * NW]
            entry = b.copy()           # entry of the left shard is empty, so exit = b
            right = emu_api.scan(pat, text, left_ctx=cut, base=cut, entry=entry)
            got = np.concatenate([left["matches"], right["matches"]])
            assert right["pending"] == 0
            assert np.array_equal(got, want), (pat, cut)
