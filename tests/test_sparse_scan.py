"""The candidate-driven ("sparse") row scan of rrx_tile.cuh (phase_a_exit / phase_a_cand) on the CPU:
the emulator runs it next to the dense row scan on every tile and flags any row where hit masks or
exit sets differ (emu_api.set_sparse(2) -> err bit 0x80); match lists, lazy starts included, are
compared with the oracle.  The same functions run on the GPU (tests -m gpu run with the sparse path on)."""
import random

import numpy as np
import pytest

import emu_api
from conftest import cpu_corpus
from roaringregex_b200.configs import CONFIGS
from test_emulator import _random_regex


@pytest.fixture(autouse=True)
def _restore_mode():
    yield
    emu_api.set_sparse(1)


def _both_modes(oracle_api, pat, text, o=None):
    o = o or oracle_api.Oracle(pat)
    want = o.lazy_matches(text)
    info = None
    # (thread-owned scan, -) / (row-bounded scan, checked row by row against the dense scan) / (row-bounded scan)
    for donly, mode in ((True, 1), (False, 2), (False, 1)):
        emu_api.set_sparse(mode)
        for reduce in (True, False):
            r = emu_api.scan(pat, text, reduce=reduce, fast=True, donly=donly)
            assert r["err"] == 0, (pat, mode, reduce, donly, hex(r["err"]))
            assert r["n"] == len(want), (pat, mode, reduce, donly, r["n"], len(want))
            assert np.array_equal(r["matches"], want), (pat, mode, reduce, donly)
            info = r["info"]
    return info


def test_cuts_of_the_configs(rrx):
    want = {1: "e", 2: "@", 3: "b", 4: "c", 5: "q"}
    for cid, byte in want.items():
        i = rrx.RRegex(CONFIGS[cid].pattern).info
        assert i["sparse_ok"] == 1 and chr(i["cand_byte"]) == byte and i["donly"] == 1, (cid, i)
    # no single byte lies on every path: dense scan only
    for pat in ("[a-z]+", "(a|b)c|d", "[ab]x"):
        i = rrx.RRegex(pat).info
        assert (i["sparse_ok"] == 0) == (pat != "[ab]x"), (pat, i)


@pytest.mark.parametrize("cfg_id", [1, 2, 3, 4, 5])
def test_config_corpora_sparse_equals_dense(oracle_api, cfg_id):
    cfg = CONFIGS[cfg_id]
    n = 150_000 if cfg_id != 4 else 50_000
    text = cpu_corpus(cfg, 0, n)
    emu_api.sparse_stats()
    _both_modes(oracle_api, cfg.pattern, text)
    emu_api.set_sparse(1)
    emu_api.sparse_stats()
    emu_api.scan(cfg.pattern, text)
    sparse_rows, dense_rows = emu_api.sparse_stats()
    assert dense_rows == 0, (sparse_rows, dense_rows)           # thread-owned scan: no dense fallback at all
    emu_api.scan(cfg.pattern, text, donly=False)
    sparse_rows, dense_rows = emu_api.sparse_stats()
    if cfg_id != 5:          # config 5 without the thread-owned scan: walks back to the last z rarely fit the row
        assert sparse_rows > 10 * dense_rows, (sparse_rows, dense_rows)


def test_known_starts_and_ambiguous_runs(oracle_api):
    """Runs fed by one injection hand their lazy start to phase C; a second candidate inside a run,
    a cut position entered from the run itself, several cut groups, and candidates at the row
    edges must all fall back to the backward scan and still give the oracle's starts."""
    rnd = random.Random(11)
    cases = [("[a-z]+@[a-z]+\\.(com|org|net)", b"ab@cd.comnetorg @@."), ("(x@y)+z", b"x@yz@ "), ("w(x@y)+z", b"wx@yz "),
             ("(a@|b)+c@", b"a@bc "), ("x@y|z@w", b"x@yzw "), ("@", b"@a"), ("@+", b"@a"), ("a@*b", b"a@b "),
             ("[ab]@[cd]", b"ab@cd"), ("z.*@b", b"z@b a"), ("(ab|cd)*e", b"abcde "), ("a{1,100}b", b"aaaaaaab ")]
    for pat, alpha in cases:
        o = oracle_api.Oracle(pat)
        for n in (1, 128, 129, 300, 4097):
            for dense_text in (False, True):
                wts = None if dense_text else [1 if ch in b"@eb" else 12 for ch in alpha]
                text = bytes(rnd.choices(alpha, weights=wts, k=n))
                _both_modes(oracle_api, pat, text, o)


def test_candidate_on_every_row_boundary(oracle_api):
    pat = "[a-z]+@[a-z]+\\.(com|org|net)"
    o = oracle_api.Oracle(pat)
    addr = b"someone@example.org"
    for edge in (128, 256, 4096):
        for shift in range(len(addr) + 2):
            text = bytearray(b" " * (edge + 200))
            text[edge - shift:edge - shift + len(addr)] = addr
            text[edge + 150:edge + 150 + 9] = b"a@b.net x"
            _both_modes(oracle_api, pat, bytes(text), o)


def test_random_regexes_with_rare_candidates(oracle_api):
    rnd = random.Random(404)
    done = capable = 0
    while done < 100:
        pat = _random_regex(rnd)
        try:
            o = oracle_api.Oracle(pat)
        except Exception:
            continue
        if o.n_raw > 300:
            continue
        done += 1
        wts = [rnd.choice([1, 1, 40, 40, 200]) for _ in range(4)]
        text = bytes(rnd.choices(b"abc\n", weights=wts, k=rnd.choice([100, 1000, 4200])))
        capable += _both_modes(oracle_api, pat, text, o)["sparse_ok"]
    assert capable > 30
