#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the scan kernel into a small JSON (the metrics DESIGN.md / VERDICT cite).
usage: summarise_ncu.py <report.ncu-rep> <out.json> [note]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, r = rows[0], rows[-1]
d = dict(zip(h, r))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "derived__smsp__sass_thread_inst_executed_op_spill", "local_load", "lts__t_sector_hit_rate.pct"]
s = {k: d[k] for k in keys if k in d}
for k, v in d.items():
    if "dram__bytes" in k and k.endswith(".sum") and k not in s:
        s[k] = v
if len(sys.argv) > 3:
    s["note"] = sys.argv[3]
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
