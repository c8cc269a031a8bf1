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
import os
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "roaringregex_b200", "csrc")
def func_ranges(path):
    """(first_line, name) of every function in a header, by a crude scan for definitions."""
    out = []
    for i, ln in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:RRX_DEV|RRX_HD|__device__ __forceinline__)\s+[\w:<>\s\*&]+?\b(\w+)\s*\(", ln)
        if m and not ln.strip().endswith(";"):
            out.append((i, m.group(1)))
    return out
TILE = func_ranges(os.path.join(SRC, "rrx_tile.cuh"))
KERN = func_ranges(os.path.join(SRC, "rrx_kernel.cuh"))
MARK = [(i, m.group(1)) for i, ln in enumerate(open(os.path.join(SRC, "rrx_kernel.cuh")), 1) for m in [re.search(r"RRX_T\((\d)\)", ln)] if m and "define" not in ln]
SEG = {0: "body.ticket", 1: "body.stage", 2: "body.phaseA-call", 3: "body.lookback-call", 4: "body.S2", 5: "body.B+scan", 6: "body.reserve", 7: "body.C", 8: "body.loop-end"}
HELPERS = {"lds_u8", "lds_tab1", "or_and", "load_B", "rotl32", "fshl", "fshr", "popc32", "ctz32", "any_w", "meets", "popc_w", "row_nbytes",
           "rel_byte", "row_byte", "chunk_ptr", "lo", "follow", "rfollow", "first", "accept", "all", "init", "h", "follow_terms", "rrx_mad",
           "comp_ident", "comp_compose", "comp_apply", "comp_closed", "load_row_comp", "store_row_comp", "store_comp_global", "load_comp_global",
           "bcast_comp", "publish_comp", "try_comp", "wait_comp", "rrx_backoff", "rrx_fence", "globaltimer_ns", "carve", "bytes"}
def owner(fr, line):
    name = None
    for l0, n in fr:
        if l0 <= line: name = n
        else: break
    return name
def phase(chain):
    if not chain: return "?"
    for f, l in chain:                        # innermost first: the first frame that is not a small helper
        if f.endswith("tile.cuh"):
            n = owner(TILE, l)
            if n and n not in HELPERS: return n
        elif f.endswith("kernel.cuh"):
            n = owner(KERN, l)
            if n and n not in HELPERS and n != "rrx_scan_body": return n
            if n == "rrx_scan_body":
                seg = 0
                for l0, k in MARK:
                    if l0 <= l: seg = int(k) + 1
                return SEG.get(seg, "body")
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

if len(sys.argv) > 3:            # per-line detail of one function: attribute_lines.py cubin rep <function>
    want = sys.argv[3]
    by_line = collections.defaultdict(lambda: [0, 0])
    for r in rows:
        off = int(r["Address"], 16) - base
        chain, sass = off2loc.get(off, (None, ""))
        if phase(chain) != want: continue
        # the frame inside the wanted function
        key = None
        for f, l in chain:
            if (f.endswith("tile.cuh") and owner(TILE, l) == want) or (f.endswith("kernel.cuh") and owner(KERN, l) == want):
                key = (f, l); break
        key = key or chain[0]
        by_line[key][0] += int(r["Instructions Executed"]); by_line[key][1] += int(r["# Samples"])
    print("---- lines of", want)
    for (f, l), (n, s_) in sorted(by_line.items(), key=lambda kv: kv[0][1]):
        if n * 500 < tot_i: continue
        print(f"{f}:{l:5d} inst {n:12d} {100*n/tot_i:5.1f}%  samples {s_:6d}")
