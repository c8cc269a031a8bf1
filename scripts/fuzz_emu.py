#!/usr/bin/env python
"""Time-bounded fuzz of the tile functions on the CPU: random regexes x random texts through the host emulator
(tests/emu: the same rrx_tile.cuh the kernels compile) against the oracle — whole texts in the thread-owned and the
row-bounded scan, and two-shard scans with entry set and left context (with the documented fallback when a
verification walk reaches the edge of the readable text).  TEST INFRASTRUCTURE: the oracle is the checker here.

    python scripts/fuzz_emu.py --seed 1 --minutes 20        # prints one line per failure, a summary at the end
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import emu_api  # noqa: E402
from oracle import api as oracle_api  # noqa: E402

ALPHABETS = [b"abc\n", b"abcxyzq ", b"ab@.com ", b"aab", b"abc" + bytes([0xe9, 0x80])]


def regex(rnd, letters, depth=0):
    ls = [chr(c) for c in letters if chr(c).isalnum()]
    atoms = ls + ["[" + "".join(rnd.sample(ls, min(2, len(ls)))) + "]", "[^" + rnd.choice(ls) + "]", ".",
                  rnd.choice(ls) + rnd.choice(ls), "(" + rnd.choice(ls) + "|" + rnd.choice(ls) + ")", ".*", "[a-z]+"]
    if depth > 3:
        return rnd.choice(atoms)
    kind = rnd.randint(0, 9)
    if kind <= 1:
        return rnd.choice(atoms)
    if kind in (2, 8):
        return regex(rnd, letters, depth + 1) + regex(rnd, letters, depth + 1)
    if kind == 3:
        return "(" + regex(rnd, letters, depth + 1) + "|" + regex(rnd, letters, depth + 1) + ")"
    if kind == 4:
        return "(" + regex(rnd, letters, depth + 1) + ")" + rnd.choice(["*", "+", "?"])
    if kind == 5:
        m = rnd.choice([1, 1, 2, 3, 7, 40])
        hi = rnd.choice(["", ",", "," + str(m + rnd.choice([1, 2, 5, 60]))])
        return "(" + regex(rnd, letters, depth + 1) + "){" + str(m) + hi + "}"
    if kind == 9:
        return regex(rnd, letters, depth + 1) + rnd.choice(ls)               # a literal tail: likely a single-byte cut
    return regex(rnd, letters, depth + 1) + rnd.choice(atoms)


def one_case(rnd, stats):
    letters = rnd.choice(ALPHABETS)
    pat = regex(rnd, letters)
    try:
        o = oracle_api.Oracle(pat)
    except Exception:
        return None
    if o.n_raw > 700:
        return None
    wts = [rnd.choice([1, 1, 5, 40, 200]) for _ in letters]
    n = rnd.choice([0, 1, 127, 128, 129, 700, 4095, 4096, 4097, 9000, 33000, 70000])
    text = bytes(rnd.choices(letters, weights=wts, k=n))
    want = o.lazy_matches(text)
    if len(want) > 200000:
        return None
    stats["cases"] += 1
    for donly in (True, False):
        for reduce in (True, False):
            r = emu_api.scan(pat, text, reduce=reduce, donly=donly)
            if r["soft"] and donly:
                stats["soft"] += 1
            if donly and r["err"] == 0x20:
                stats["budget_fallback"] += 1       # a verification walk out of budget: the product redoes the scan row-bounded
                continue
            if r["err"] or r["n"] != len(want) or not np.array_equal(r["matches"], want):
                return ("whole", pat, letters, wts, n, donly, reduce, hex(r["err"]), r["n"], len(want))
    if n >= 700:
        cut = rnd.choice([n // 2, (n // 2) & ~127, ((n // 2) & ~4095) or 128, n - 1, 1])
        cut = min(max(cut, 1), n - 1)
        ctx = rnd.choice([0, min(cut, 128), cut])
        donly = True
        left = emu_api.scan(pat, text[:cut])
        nw = left["info"]["NW"]
        right = emu_api.scan(pat, text[cut - ctx:], left_ctx=ctx, base=cut, entry=left["shard_comp"][1:1 + nw].copy())
        if (left["err"] | right["err"]) & 0x20:
            left["soft"], left["err"], right["err"] = 16, left["err"] & ~0x20, right["err"] & ~0x20
        if left["soft"] and left["err"] == 0:
            left = emu_api.scan(pat, text[:cut], donly=False)
            donly = False
        if (right["soft"] and ctx < cut) or not donly:
            stats["shard_fallback"] += 1
            if donly:
                left = emu_api.scan(pat, text[:cut], donly=False)
            right = emu_api.scan(pat, text[cut - ctx:], left_ctx=ctx, base=cut, entry=left["shard_comp"][1:1 + nw].copy(), donly=False)
        if right["err"] == 2 and not left["err"]:
            stats["pending_overflow"] += 1          # more than the emulator's 64 pending-start slots: a documented, loud limit
            return None
        if left["err"] or right["err"]:
            return ("shard-err", pat, letters, wts, n, cut, ctx, hex(left["err"]), hex(right["err"]))
        got = np.concatenate([left["matches"], right["matches"]])
        stats["shards"] += 1
        if len(got) != len(want) or not np.array_equal(got[:, 1], want[:, 1]):
            return ("shard-ends", pat, letters, wts, n, cut, ctx, len(got), len(want))
        if right["pending"] == 0 and not np.array_equal(got, want):
            return ("shard-starts", pat, letters, wts, n, cut, ctx)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=5.0)
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    stats = dict(cases=0, shards=0, soft=0, shard_fallback=0, pending_overflow=0, budget_fallback=0, failures=0)
    t_end = time.time() + 60 * a.minutes
    it = 0
    while time.time() < t_end:
        it += 1
        state = rnd.getstate()
        bad = one_case(rnd, stats)
        if bad:
            stats["failures"] += 1
            print("FAIL", "seed", a.seed, "iter", it, bad, flush=True)
            if stats["failures"] >= 20:
                break
        del state
    print("fuzz_emu seed", a.seed, stats, flush=True)
    return 1 if stats["failures"] else 0


if __name__ == "__main__":
    sys.exit(main())
