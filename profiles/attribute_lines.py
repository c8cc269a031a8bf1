#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an ncu report to source phases.
usage: attribute_lines.py <cubin> <ncu-rep> [kernel-name]
Joins `nvdisasm -gi` (offset -> file:line with inline chain) with `ncu --page source --csv`
(address -> instructions executed, stall samples)."""
import csv, io, re, subprocess, sys, collections

cubin, rep = sys.argv[1], sys.argv[2]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
off2loc, cur, pend = {}, None, []
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if pend:
            cur, pend = pend, []
        off2loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[hdr_i:]))))
base = int(rows[0]["Address"], 16)
PH = [("kernel.cuh", 98, 176, "lookback"), ("tile.cuh", 401, 546, "A"), ("tile.cuh", 547, 577, "A2"),
      ("tile.cuh", 620, 737, "S"), ("tile.cuh", 738, 781, "B"), ("tile.cuh", 782, 837, "C.resolve"), ("tile.cuh", 838, 876, "C.list")]
def phase(chain):
    if not chain: return "?"
    for f, l in chain:                        # innermost first
        for pf, lo, hi, name in PH:
            if f.endswith(pf) and lo <= l <= hi: return name
    for f, l in chain:
        if f.endswith("kernel.cuh"):
            if 232 <= l <= 241: return "ticket/prefetch"
            if 242 <= l <= 267: return "stage"
            if 290 <= l <= 308: return "count-scan/seg"
            return "body:%d" % l
    return chain[-1][0] + ":%d" % chain[-1][1]
tot_i = tot_s = 0
acc = collections.defaultdict(lambda: [0, 0])
for r in rows:
    off = int(r["Address"], 16) - base
    chain, sass = off2loc.get(off, (None, ""))
    n, s = int(r["Instructions Executed"]), int(r["# Samples"])
    tot_i += n; tot_s += s
    p = phase(chain)
    acc[p][0] += n; acc[p][1] += s
print("total warp instr", tot_i, "samples", tot_s)
for p, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    if n * 1000 < tot_i and s * 1000 < tot_s: continue
    print(f"{p:22s} inst {n:12d} {100*n/tot_i:5.1f}%   samples {s:7d} {100*s/max(tot_s,1):5.1f}%")
