"""Print the essentials of a bench.py JSON line (headline, parity, sub-records)."""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"headline {l['value']:.0f} GB/s  step {l['ms_per_step']:.4f} ms  kernel {l['roofline']['kernel_ms']:.4f} ms  frac {l['roofline']['frac']:.4f}  n_gpus {l['n_gpus']}")
if l.get("e2e"):
    print("e2e", round(l["e2e"]["value"], 1), "pageable", (l["e2e"].get("pageable") or {}).get("value"))
if l.get("parity"):
    print("parity ok:", l["parity"]["ok"])
for k, v in (l.get("configs") or {}).items():
    if "error" in v:
        print(k, v)
        continue
    fr = v["roofline"]["frac"] if "roofline" in v else v.get("roofline_frac_per_gpu")
    print(f"{k:16s} {v['value']:8.1f} GB/s  kernel_ms {v.get('kernel_ms', 0):9.3f}  frac {fr:.4f}  matches {v['matches']}  parity {v['parity']['ok']}  {v.get('kernel_variant', '')}")
