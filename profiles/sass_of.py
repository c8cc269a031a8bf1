#!/usr/bin/env python
"""Print the SASS of a cubin whose line info points into [lo,hi] of a source file (any inline frame).
usage: sass_of.py <cubin> <file-suffix> <lo> <hi>"""
import re, subprocess, sys
cubin, suffix, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
pend, cur, n = [], None, 0
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    if ln.strip().startswith(".L_"):
        print(ln.strip()); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if pend: cur, pend = pend, []
        fr = [l for f, l in (cur or []) if f.endswith(suffix) and lo <= l <= hi]
        if fr:
            n += 1
            print(f"  {m.group(1)} L{fr[-1]:<5d} {m.group(2)}")
print("instructions:", n)
