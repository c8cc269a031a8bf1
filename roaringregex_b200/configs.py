"""The five benchmark configurations of BASELINE.json / SURVEY.md §8d: regex, corpus recipe, size.
Pure data, shared by bench.py and the tests; the generators are oracle/corpus.c (CPU) and
csrc/rrx_corpus.cu (device)."""
from __future__ import annotations

from dataclasses import dataclass


def _alpha(spec) -> bytes:
    out = bytearray()
    for ch, count in spec:
        out += bytes([ord(ch)]) * count
    assert len(out) == 256, len(out)
    return bytes(out)


_OTHER = [c for c in "abcdefghijklmnoprstuvw"]            # lower-case letters except q x y z: 22 letters share 108
_C5 = [("x", 64), ("y", 64), ("z", 2), ("q", 1), (" ", 16), ("\n", 1)] + \
      [(c, 5) for c in _OTHER[:20]] + [(c, 4) for c in _OTHER[20:]]


@dataclass(frozen=True)
class Config:
    name: str
    pattern: str
    length: int
    seed: int
    mode: str            # "iid" | "records"
    alpha: bytes | None
    note: str


CONFIGS = {
    1: Config("config1", "(ab|cd)*e", 1 << 20, 1, "iid",
              _alpha([("a", 60), ("b", 60), ("c", 60), ("d", 60), ("e", 8), (" ", 8)]),
              "test_regex analogue; <=64-state path"),
    2: Config("config2", "[a-z]+@[a-z]+\\.(com|org|net)", 1 << 30, 2, "records", None,
              "1 GiB synthetic log records; <=64-state register path, HBM-bound"),
    # config 2 on unpadded, variable-length (39..254 B), unaligned lines: addresses straddle the scanner's rows
    "2v": Config("config2v", "[a-z]+@[a-z]+\\.(com|org|net)", 1 << 30, 2, "lines", None,
                 "config 2's regex over 1 GiB of variable-length unaligned log lines (robustness variant)"),
    3: Config("config3", "a{1,100}b", 1 << 32, 3, "iid",
              _alpha([("a", 240), ("b", 1), (" ", 8), ("\n", 1), ("c", 6)]),
              "<=256-state path (flattened N=102)"),
    4: Config("config4", "(a|b){1,300}c", 1 << 33, 4, "iid",
              _alpha([("a", 120), ("b", 120), ("c", 1), (" ", 8), ("\n", 1), ("d", 6)]),
              ">256-state path (flattened N=602)"),
    5: Config("config5", "(x|y)*z.*q", 1 << 36, 5, "iid", _alpha(_C5),
              "64 GiB over 1/2/4/8 GPUs; unbounded carry ('.' matches '\\n')"),
}
