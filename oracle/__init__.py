"""TEST INFRASTRUCTURE ONLY — CPU oracle for the RoaringRegex hot path.

ref_nfa.py     restatement of the reference's ERE -> raw NFA construction (Parser.cpp, NFA.cc)
lazy_oracle.c  restatement of the reference's per-byte step + flags and the LAZY-v1 enumerator
corpus.c       synthetic corpus generator, CPU side
_ref/          the UNMODIFIED reference compiled from /root/reference (whole-string acceptance)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package.  Nothing under roaringregex_b200/ does.
"""
