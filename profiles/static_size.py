#!/usr/bin/env python
"""Static SASS instruction count of a cubin per source function (innermost non-helper frame), to see what the
per-tile hot code weighs in the instruction cache.  usage: static_size.py <cubin>"""
import collections
import re
import subprocess
import sys

dis = subprocess.run(["nvdisasm", "-gi", "-c", sys.argv[1]], capture_output=True, text=True).stdout
cur, pend, acc, tot = None, [], collections.Counter(), 0
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        if pend:
            cur, pend = pend, []
        tot += 1
        acc[cur[0] if cur else ("?", 0)] += 1
src = {}
def fn_of(f, l):
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "roaringregex_b200", "csrc", f)
    if f not in src:
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    name = "?"
    for i in range(min(l, len(src[f])) - 1, -1, -1):
        m = re.match(r"\s*(?:RRX_DEV|RRX_HD|RRX_COLD|__device__ __forceinline__)\s+[\w:<>\s\*&]+?\b(\w+)\s*\(", src[f][i])
        if m:
            name = m.group(1)
            break
    return name
byfn = collections.Counter()
for (f, l), n in acc.items():
    byfn[fn_of(f, l)] += n
print("total SASS instructions", tot, "=", tot * 16 // 1024, "KiB")
for k, n in byfn.most_common(30):
    print(f"{k:28s} {n:6d}  {n * 16 / 1024:6.1f} KiB")
