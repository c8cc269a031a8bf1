"""Regenerates tests/golden/*.json from the compiled, unmodified reference (oracle/_ref).
Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

  nfa_dumps.json    the reference's own NFA::print() output (NFA.cc:14-41), parsed: raw state
                    count, final states, forward and backward rows  
  acceptance.json   whole-string acceptance answers of the reference's AcceptanceIterator
                    (regex.h:150-165) on fixed and seeded-random strings (Appendix B.2)
  lazy_v1.json      LAZY-v1 match lists obtained by BRUTE FORCE over all (start,end) pairs with the
                    reference's acceptance as the only membership test (Appendix B.3 / D.1)
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import api  # noqa: E402

DUMP_PATTERNS = ["ab", "a|b", "a*", "a+", "a?b", "a{2,3}", "a{0,1}b", "x*|y", "[a-c]+d", "(ab|cd)*e", "(x|y)*z.*q",
                 "[a-z]+@[a-z]+\\.(com|org|net)", "a{3,}", "[^a]", "[\\]]", "ab$", "^ab", "a\\.b\\*", "(a|bc)+d?e{2}",
                 "a{1,20}b", "a{1,40}b", "(a|b){1,40}c", "a{1,80}b", ".*a.{3}", "(a|b)*a(a|b){12}"]

ACCEPT_FIXED = {
    "(ab|cd)*e": ["ababcde", "e", "xxababe", "abex"],
    "x*|y": ["xy", "x", "y", "", "xxy", "yx"],
    "a{0,2}b": ["b", "ab", "aaab", "aaaab"],
    "a{2,3}b": ["aaab", "aab", "ab"],
    "a{3,}": ["aaaa", "aa", "aaa"],
    "[^a]": ["^", "b", "a"],
    "[\\]]": ["\\", "]", "a"],
    "ab$": ["ab"], "^ab": ["ab"],
    "(x|y)*z.*q": ["xyzaaq", "xyzaaqx", "zq", "z\nq"],
    "[a-z]+@[a-z]+\\.(com|org|net)": ["foo@bar.com", "foo@bar.co", "a@b.org", "a@b.ne."],
    "a{1,80}b": ["a" * 80 + "b", "a" * 81 + "b", "b", "ab"],
    "(a|b){1,40}c": ["c", "ab" * 20 + "c", "ab" * 20 + "ac"],
}
ACCEPT_RANDOM = [("(ab|cd)*e", "abcde"), ("x*|y", "xy"), ("a{0,2}b", "ab"), ("[a-c]+d", "abcd"), ("(x|y)*z.*q", "xyzqa"),
                 ("a?b|c{2,3}", "abc"), ("(a|bc)+d?e{2}", "abcde"), ("[a-z]+@[a-z]+\\.(com|org|net)", "acomnetrg@."),
                 ("a{1,40}b", "ab"), ("(a|b){1,40}c", "abc"), ("a{1,80}b", "ab"), (".*a.{3}", "ab"), ("(a|b)*a(a|b){12}", "ab")]
LAZY_CASES = [("(ab|cd)*e", "xxababe cde e"), ("[a-z]+@[a-z]+\\.(com|org|net)", "foo@bar.com x a@b.org c@d.comet@e.net"),
              ("a{1,100}b", None), ("(a|b){1,300}c", None),     # > 256 raw states: not runnable on the reference
              ("(x|y)*z.*q", "xyzaaqbbq q zq"), ("ab|bc", "abc"), ("a+", "aaa"), ("a{1,40}b", "aaab b ab aab"),
              ("(a|b){1,40}c", "abbac c bc"), ("x*|y", "xxyy xy"), ("a{0,2}b", "b ab aaab aaaab")]
LAZY_RANDOM = [("(ab|cd)*e", "abcde "), ("x*|y", "xyz"), ("a{0,2}b", "ab "), ("[a-c]+d", "abcd "), ("(x|y)*z.*q", "xyzq a"),
               ("a?b|c{2,3}", "abc"), ("(a|bc)+d?e{2}", "abcde"), ("a{1,20}b", "ab c"), (".*a.{3}", "ab")]


def brute(pat, t: bytes):
    out = []
    for e in range(1, len(t) + 1):
        for s in range(e - 1, -1, -1):
            if api.ref_accepts(pat, t[s:e]):
                out.append([s, e])
                break
    return out


def main():
    assert api.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    dumps = {}
    for p in DUMP_PATTERNS:
        n, fin, fwd, bwd = api.ref_dump(p)
        dumps[p] = dict(n=n, finals=fin, fwd={str(s): {str(c): t for c, t in r.items()} for s, r in fwd.items()},
                        bwd={str(s): {str(c): t for c, t in r.items()} for s, r in bwd.items()})
    json.dump(dumps, open(os.path.join(HERE, "nfa_dumps.json"), "w"), indent=0, sort_keys=True)

    rnd = random.Random(20261018)
    acc = []
    for p, texts in ACCEPT_FIXED.items():
        for t in texts:
            acc.append([p, t, int(api.ref_accepts(p, t.encode("latin-1")))])
    for p, alpha in ACCEPT_RANDOM:
        for _ in range(60):
            n = rnd.randint(0, 100 if "{1," in p else 14)
            t = "".join(rnd.choice(alpha) for _ in range(n))
            acc.append([p, t, int(api.ref_accepts(p, t.encode("latin-1")))])
    json.dump(acc, open(os.path.join(HERE, "acceptance.json"), "w"), indent=0)

    lazy = []
    for p, t in LAZY_CASES:
        if t is not None:
            lazy.append([p, t, brute(p, t.encode("latin-1"))])
    for p, alpha in LAZY_RANDOM:
        for _ in range(12):
            t = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 16)))
            lazy.append([p, t, brute(p, t.encode("latin-1"))])
    json.dump(lazy, open(os.path.join(HERE, "lazy_v1.json"), "w"), indent=0)
    print(len(dumps), "dumps,", len(acc), "acceptance answers,", len(lazy), "lazy vectors")


if __name__ == "__main__":
    main()
