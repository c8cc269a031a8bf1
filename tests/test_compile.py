"""Host side of the product on CPU: librrx.so loads and exports what include/rrx.h declares, the
host compiler reproduces the reference's raw tables (golden dumps) and flattens them exactly, and
every scan entry point fails loudly when there is no GPU (no CPU fallback).  No compute calls."""
import json
import os
import random
import re

import numpy as np
import pytest

from oracle import ref_nfa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMPS = json.load(open(os.path.join(ROOT, "tests", "golden", "nfa_dumps.json")))


def test_library_exports_every_declared_symbol(rrx):
    hdr = open(os.path.join(ROOT, "include", "rrx.h")).read()
    names = sorted(set(re.findall(r"\b(rrx_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = rrx.lib()
    for n in names:
        assert hasattr(L, n), f"librrx.so does not export {n}"
    assert b"sm_100a" in L.rrx_version()


@pytest.mark.parametrize("pattern", sorted(DUMPS))
def test_host_compiler_reproduces_reference_tables(rrx, pattern):
    """The atom-tracking restatement rebuilds the reference's raw forward rows (Appendix B.1)."""
    g = DUMPS[pattern]
    r = rrx.RRegex(pattern)
    assert r.info["n_raw"] == g["n"]
    assert r.raw_finals() == g["finals"]
    fwd = {int(s): {int(c): t for c, t in row.items()} for s, row in g["fwd"].items()}
    for s in range(g["n"]):
        for c in range(128):
            assert r.raw_row(s, c) == fwd.get(s, {}).get(c, []), (pattern, s, c)


def test_config_shapes(rrx):
    # raw n / flattened N of SURVEY.md §0.5 and the reference's width class (Parser.cpp:165-168)
    want = {"(ab|cd)*e": (10, 6, 0), "[a-z]+@[a-z]+\\.(com|org|net)": (30, 16, 0), "a{1,100}b": (301, 102, 3),
            "(a|b){1,300}c": (1501, 602, 3), "(x|y)*z.*q": (10, 6, 0)}
    for pat, (n, nf, wc) in want.items():
        i = rrx.RRegex(pat).info
        assert (i["n_raw"], i["n_flat"], i["ref_width_class"]) == (n, nf, wc), pat
        j = rrx.RRegex(pat, rrx.COMPILE_NO_REDUCE).info
        assert j["n_device"] == nf - 1


def _flat_sim(r, text):
    """Simulate the flattened automaton (first/follow/classes) in numpy-free Python: yields the
    set of RAW states after each byte as the union of post(k) ... here compared through accept."""
    P = r.info["n_flat"] - 1
    W = (P + 31) // 32
    first = int.from_bytes(r.flat_table(0).tobytes(), "little")
    accept = int.from_bytes(r.flat_table(1).tobytes(), "little")
    fol = r.flat_table(2).reshape(P, W)
    follow = [int.from_bytes(fol[k].tobytes(), "little") for k in range(P)]
    cls = r.flat_table(3).reshape(P, 4)
    B = [0] * 256
    for k in range(P):
        v = int.from_bytes(cls[k].tobytes(), "little")
        for c in range(128):
            if v >> c & 1:
                B[c] |= 1 << k
    return first, accept, follow, B


@pytest.mark.parametrize("pattern,alpha", [("(ab|cd)*e", "abcde"), ("x*|y", "xy"), ("a{0,2}b", "ab"), ("[^a]b", "ab^"),
                                           ("(x|y)*z.*q", "xyzq\n"), ("(a|bc)+d?e{2}", "abcde"), ("a{1,20}b", "ab"),
                                           ("[a-z]+@[a-z]+\\.(com|org|net)", "acomnetrg@.")])
def test_flattening_preserves_acceptance_at_every_step(rrx, pattern, alpha):
    """G' = follow(G) & B[c] on positions accepts exactly when the raw automaton does,
    after every prefix (anchored run, no re-injection)."""
    r = rrx.RRegex(pattern, rrx.COMPILE_NO_REDUCE)
    raw = ref_nfa.compile_raw(pattern)
    first, accept, follow, B = _flat_sim(r, None)
    rnd = random.Random(3)
    for _ in range(120):
        t = [ord(rnd.choice(alpha)) for _ in range(rnd.randint(0, 30))]
        cur_raw = 1 << raw.init
        G, started = 0, False
        for c in t:
            nxt = 0
            for i in ref_nfa._bits(cur_raw):
                nxt |= raw.fwd.get((i, c), 0)
            cur_raw = nxt
            f = first if not started else 0
            for k in ref_nfa._bits(G):
                f |= follow[k]
            G, started = f & B[c], True
            assert bool(G & accept) == bool(cur_raw & raw.finals), (pattern, bytes(t))
            assert bool(G) == bool(cur_raw)


def test_parse_errors(rrx):
    for bad in ["[", "*a", "a)", "(|a)", "\x80"]:
        with pytest.raises(rrx.InvalidExpression):
            rrx.RRegex(bad.encode("latin-1"))


def test_device_program_is_self_consistent(rrx):
    """Terms + bicliques + rows of the device program cover exactly the pruned follow relation."""
    from tests_support import decode_blob, apply_follow
    for pat in ["(ab|cd)*e", "[a-z]+@[a-z]+\\.(com|org|net)", "a{1,100}b", "(a|b){1,300}c", "(x|y)*z.*q",
                "(a|bc)+d?e{2}", "(ab|ba|aab|bba|abab)*c", "a(b|c|d|e|f|g|h|i|j|k)(l|m)z{1,40}q"]:
        for flags in (rrx.COMPILE_DEFAULT, rrx.COMPILE_NO_REDUCE):
            prog = decode_blob(rrx.RRegex(pat, flags).blob())
            for rev in (False, True):
                rows = prog["rfollow" if rev else "follow"]
                for k in range(prog["P"]):
                    assert apply_follow(prog, 1 << k, rev) == rows[k], (pat, flags, rev, k)


def test_scan_without_gpu_fails_loudly(rrx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = rrx.RRegex("(ab|cd)*e")
    with pytest.raises(rrx.RrxError) as e:
        r.scan_host(b"xxababe cde e")
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(rrx.RrxError):
        rrx.Scanner(r, 1 << 20, 1024)


def test_specialised_kernel_compiles_without_gpu(rrx, tmp_path):
    """NVRTC turns the generated policy + the embedded kernel headers into an sm_100a cubin."""
    r = rrx.RRegex("[a-z]+@[a-z]+\\.(com|org|net)")
    src = r.jit_source()
    assert "struct ProgJit" in src and "rrx_mad" in src
    out = tmp_path / "k.cubin"
    try:
        r.jit_compile(str(out))
    except rrx.RrxError as e:
        if "NVRTC could not be loaded" in str(e):
            pytest.skip("no NVRTC on this machine")
        raise
    assert out.stat().st_size > 10000
    # multi-word automata are specialised too (word-by-word funnel shifts, zero words dropped)
    for pat, words in (("a{1,100}b", 4), ("(a|b){1,300}c", 16)):
        rw = rrx.RRegex(pat)
        assert rw.info["words"] == words and "__funnelshift_l" in rw.jit_source()
        rw.jit_compile(str(out))
        assert out.stat().st_size > 10000


def test_both_kernel_variants_compile_for_every_config(rrx, tmp_path, monkeypatch):
    """Thread-owned programs have a lean and a straddle kernel (rrx_jit.h): both must compile for sm_100a."""
    from roaringregex_b200.configs import CONFIGS
    sizes = {}
    for variant in ("0", "1"):
        monkeypatch.setenv("RRX_JIT_VARIANT", variant)
        for cid in (1, 2, 3, 4, 5):
            r = rrx.RRegex(CONFIGS[cid].pattern)
            assert r.info["donly"] == 1
            out = tmp_path / f"k{cid}_{variant}.cubin"
            try:
                r.jit_compile(str(out))
            except rrx.RrxError as e:
                if "NVRTC could not be loaded" in str(e):
                    pytest.skip("no NVRTC on this machine")
                raise
            sizes[(cid, variant)] = out.stat().st_size
    assert all(v > 10000 for v in sizes.values())
    assert any(sizes[(c, "0")] != sizes[(c, "1")] for c in (1, 2, 3, 4, 5))      # the variants are different kernels
