"""Reverse-walk skip state of the lazy start scan (rrx_compile.cc find_reverse_skip, rrx_tile.cuh
resolve_and_write): patterns with an absorbing `.*`-like state let the backward walk run over the text
word by word; results must not change — whole texts, and shards with left context / entry / hand-over."""
import random

import numpy as np

import emu_api


def test_skip_state_found_where_expected(rrx):
    def rs(pat):
        src = rrx.RRegex(pat).jit_source()
        n = int(src.split("rs_n(const ProgView&) const { return ")[1].split("u;")[0])
        b = int(src.split("rs_bytes(const ProgView&) const { return ")[1].split("u;")[0], 16)
        return n, chr(b & 0xff)
    assert rs("(x|y)*z.*q") == (1, "z")
    assert rs("z.*@b") == (1, "z")
    assert rs("[a-z]+@[a-z]+\\.(com|org|net)")[0] == 0
    assert rs("a{1,100}b")[0] == 0


def test_walks_with_skip_state_match_the_oracle(oracle_api):
    rnd = random.Random(7)
    pats = ["(x|y)*z.*q", "z.*@b", "a.*b.*c", "a.*b", "(a|b).*c", "a[^b]*c", "(xy)*z.*q.*w", "a(.|b)*c"]
    for it in range(120):
        pat = rnd.choice(pats)
        alpha = rnd.choice([b"xyzq a", b"abc\n ", b"z@b a", b"xyzqw ", b"abc" + bytes([200])])
        wts = [rnd.choice([1, 1, 30, 100]) for _ in alpha]
        n = rnd.choice([1, 100, 129, 300, 4096, 4097, 9000])
        text = bytes(rnd.choices(alpha, weights=wts, k=n))
        want = oracle_api.Oracle(pat).lazy_matches(text)
        for fast in (True, False):
            r = emu_api.scan(pat, text, fast=fast)
            assert r["err"] == 0 and r["n"] == len(want) and np.array_equal(r["matches"], want), (pat, n, fast)
        if n > 600:                          # a shard with entry set and some (or no) left context
            cut = (n // 2) & ~127
            left = emu_api.scan(pat, text[:cut])
            nw = left["info"]["NW"]
            entry = left["shard_comp"][1:1 + nw].copy()
            ctx = rnd.choice([0, 128, cut])
            right = emu_api.scan(pat, text[cut - ctx:], left_ctx=ctx, base=cut, entry=entry)
            if right["soft"] and ctx < cut:
                # thread-owned scan: a verification walk reached the left edge of the readable text although there is
                # text further left — both shards are redone with the row-bounded scan, as sharded.py does
                left = emu_api.scan(pat, text[:cut], donly=False)
                entry = left["shard_comp"][1:1 + nw].copy()
                right = emu_api.scan(pat, text[cut - ctx:], left_ctx=ctx, base=cut, entry=entry, donly=False)
            got = np.concatenate([left["matches"], right["matches"]])
            assert np.array_equal(got[:, 1], want[:, 1]), (pat, n, cut, ctx)
            if right["pending"] == 0:
                assert np.array_equal(got, want), (pat, n, cut, ctx)
