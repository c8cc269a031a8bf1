"""scripts/fuzz_emu.py stays runnable and clean: a few dozen cases of the random regex x text x shard-cut fuzz (the long
runs are recorded in profiles/r02_fuzz.txt)."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def test_fuzz_cases_are_clean():
    import fuzz_emu
    rnd = random.Random(2027)
    stats = dict(cases=0, shards=0, soft=0, shard_fallback=0, pending_overflow=0, budget_fallback=0, failures=0)
    bad = []
    while stats["cases"] < 40:
        r = fuzz_emu.one_case(rnd, stats)
        if r:
            bad.append(r)
    assert not bad, bad[:3]
    assert stats["shards"] > 5
