#!/usr/bin/env python
"""bench.py — scanned GB/s with bit-exact lazy matches (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--size BYTES] [--impl ours|reference]

A "step" is one pass of the hot path (regex scan -> LAZY-v1 match list) over one synthetic corpus.
N=1 workload: config 2 of BASELINE.json (e-mail regex over 1 GiB of synthetic log records).
For N>1 (torchrun) the corpus is `--size` bytes PER RANK (weak scaling), sharded into contiguous
byte ranges, one per GPU.

value  : corpus resident in HBM, program compiled, CUDA events on the launching stream
e2e    : the same scan through the C ABI with HOST buffers (pinned), H2D of the corpus and D2H of
         the match list inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--size", type=int, default=0, help="corpus bytes per rank (default: the config's size, capped for N>1)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(cfg, sample_per_thread: int):
    """The oracle port (reference algorithm: per-active-state row OR, injection, flag test, backward
    start scan) on all host cores, each core scanning its own contiguous piece of the corpus."""
    from oracle import api as oracle_api
    oracle_api.build(ref=False)
    cores = os.cpu_count() or 1
    o = oracle_api.Oracle(cfg.pattern)
    n = sample_per_thread * cores
    text = oracle_api.corpus_iid(cfg.seed, cfg.alpha, 0, n) if cfg.mode == "iid" else oracle_api.corpus_records(cfg.seed, 0, n)
    o.time_parallel(text, min(sample_per_thread, 1 << 20), cores)         # warm
    sec, m = o.time_parallel(text, sample_per_thread, cores)
    out = {"value": n / sec / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
           "sample": f"first {n >> 20} MiB of the corpus, {sample_per_thread >> 20} MiB per thread, {cores} threads, full LAZY-v1 match lists ({m} matches)"}
    # the compiled, unmodified reference can only do whole-string acceptance (its lazy iterator is
    # missing, src/Iterator.cpp): time that too, 1 core (it is single-threaded), for orientation.
    if oracle_api.have_ref() and o.n_raw <= 256:
        k = min(n, 16 << 20)
        t = bytes(text[:k]).replace(b"\0", b" ")
        sec_r, _ = oracle_api.ref_time_acceptance(cfg.pattern, t, 3)
        out["reference_acceptance_only"] = {"value": k / sec_r / 1e9, "unit": "GB/s", "cores": 1,
                                            "note": "oracle/_ref AcceptanceIterator, anchored: set empties after a few bytes (lower bound on cost)"}
    return out


def run_reference(args, cfg):
    """Reference arm: the reference's CPU algorithm for this path on all host cores.  The compiled
    reference (oracle/_ref) has no match enumerator (src/Iterator.cpp is empty), so the timed code is
    the oracle port of its algorithm (kind "port"); each step scans a bounded sample of the same
    corpus (32 MiB per thread).  Under torchrun only rank 0 runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_thread = 32 << 20
    vals, res = [], None
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(cfg, per_thread)
        if i >= args.warmup:
            vals.append(res["value"])
        if time.perf_counter() - t_start > 150 and vals:        # keep the whole run within a few minutes
            break
    val = float(np.mean(vals))
    nbytes = per_thread * res["cores"]
    line = {"impl": "reference", "metric": "scanned GB/s, bit-exact lazy matches", "value": val, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": nbytes / val / 1e6,
            "higher_is_better": True, "scaling": "weak",
            "dtype": "u64 bitsets", "data": "synthetic", "vs_baseline": None,
            "config": {"workload": f"{cfg.name}: regex {cfg.pattern!r}; each step = the first {nbytes >> 20} MiB of the corpus "
                                   f"({per_thread >> 20} MiB per host thread), full LAZY-v1 match lists"},
            "cpu_baseline": {**res, "value": val},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measured_traffic(cfg_name: str, size: int):
    """DRAM bytes per launch of the scan kernel from the committed ncu --set full capture, if there is
    one for exactly this workload (profiles/measured_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "measured_traffic.json")
    if not os.path.exists(p):
        return None
    rec = json.load(open(p)).get(f"{cfg_name}:{size}")
    return rec["dram_bytes_per_launch"] if rec else None


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from roaringregex_b200.configs import CONFIGS
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import roaringregex_b200 as rrx
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # config 2 (the headline): the config's size PER GPU (weak scaling); config 5 is defined as one
    # 64 GiB corpus sharded over the GPUs (strong scaling)
    strong = args.config == 5 and not args.size
    size = args.size or (cfg.length // world if strong else min(cfg.length, 8 << 30))
    size = min(size, 120 << 30) & ~4095
    regex = rrx.RRegex(cfg.pattern)
    info = regex.info
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    # corpus of this rank, generated in place: [rank*size, (rank+1)*size) of the global text, plus
    # `halo` bytes of left context for lazy start scans that cross the shard's left edge
    off = rank * size
    halo = min(off, 1 << 20)
    d_buf = torch.empty(halo + size + 256, dtype=torch.uint8, device="cuda")
    d_text = d_buf[halo:]
    assert d_text.data_ptr() % 16 == 0
    if cfg.mode == "iid":
        rrx.corpus_iid(d_buf.data_ptr(), off - halo, halo + size, cfg.seed, cfg.alpha, sh)
    else:
        rrx.corpus_records(d_buf.data_ptr(), off - halo, halo + size, cfg.seed, sh)
    torch.cuda.synchronize()

    cap = max(1 << 20, size // 48)
    if world == 1:
        sc = rrx.Scanner(regex, size, cap)
        step = lambda: sc.scan(d_text.data_ptr(), size, off, 0, None, sh)
        for _ in range(max(args.warmup, 3)):
            step()
        res = sc.result(sh)
        assert res["error_bits"] == 0, res
        if res["n_matches"] > cap:
            cap = res["n_matches"] + 1024
            sc = rrx.Scanner(regex, size, cap)
            step()
            res = sc.result(sh)
        n_matches = res["n_matches"]
        launches_per_step = res["launches"]
        is_jit = sc.is_jit
    else:
        from roaringregex_b200 import sharded
        comm = sharded.TorchComm(dist, torch.device("cuda", local))
        probe = rrx.Scanner(regex, size, cap)
        probe.scan(d_text.data_ptr(), size, off, halo, None, sh)
        cap = max(cap, probe.result(sh)["n_matches"] + (1 << 16))
        del probe
        be = sharded.DeviceBackend(regex, d_text.data_ptr(), size, off, halo, cap, sh)
        last = {}

        def step():
            last["res"] = sharded.scan_sharded(regex, be, comm, rank, world, fetch=False)
        for _ in range(max(args.warmup, 4)):
            step()
        n_matches = last["res"].count
        launches_per_step = 4 * (1 + (2 if last["res"].rescans == 1 else last["res"].rescans))
        is_jit = be.scanner.is_jit

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scan_kernel_ms = []
    if world == 1:
        timed_scanner = sc
        sc.enable_timing(True)
    else:
        for s_ in be.scanners:
            if s_ is not None:
                s_.enable_timing(True)
    barrier()
    w0 = time.perf_counter()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    if world == 1:
        for a, b in evs:
            a.record(stream)
            step()
            b.record(stream)
    else:
        # Steps are software-pipelined: step k+1 is enqueued (scan kernels, message copy, all-gather,
        # D2H) before the host waits for step k's gathered messages, folds the entry sets and validates
        # the speculative scan.  Every step's collective and validation complete inside the timed region.
        prev = None
        for _ in range(args.steps):
            tok = sharded.scan_sharded_begin(regex, be, comm)
            if prev is not None:
                last["res"] = sharded.scan_sharded_finish(regex, be, comm, rank, world, prev, fetch=False)
                scan_kernel_ms.append(prev["scanner"].kernel_ms())
            prev = tok
        last["res"] = sharded.scan_sharded_finish(regex, be, comm, rank, world, prev, fetch=False)
        scan_kernel_ms.append(prev["scanner"].kernel_ms())
        assert last["res"].count == n_matches
    t_all1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    # N=1: device time between the first launch and the last kernel's end (CUDA events on the launching
    # stream); N>1: a step contains host-side collectives, so the barrier-to-barrier wall clock counts.
    total_ms = t_all0.elapsed_time(t_all1) if world == 1 else wall_ms
    kern_ms = [a.elapsed_time(b) for a, b in evs] if world == 1 else [wall_ms / args.steps]
    if world == 1:
        # the dominant kernel alone, timed live by CUDA events around its launch (last step; the steps are
        # identical).  Re-run one step with nothing queued behind it so the events bracket only that kernel.
        step()
        torch.cuda.synchronize()
        scan_kernel_ms = [timed_scanner.kernel_ms()]
        for _ in range(4):
            step()
            torch.cuda.synchronize()
            scan_kernel_ms.append(timed_scanner.kernel_ms())
    clocks = sampler.stop() if sampler else None
    if world == 1:
        res = sc.result(sh)
        assert res["n_matches"] == n_matches and res["error_bits"] == 0

    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        m = torch.tensor([n_matches], device="cuda", dtype=torch.int64)
        dist.all_reduce(m)
        total_matches = int(m.item())
    else:
        total_matches = n_matches
    ms_per_step = total_ms / args.steps
    total_bytes = size * world
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----------------------
    e2e = None
    if not args.no_e2e:
        e_size = min(size, 1 << 30)
        h_text = torch.empty(e_size, dtype=torch.uint8).pin_memory()
        h_text.copy_(d_text[:e_size])
        torch.cuda.synchronize()
        host_np = h_text.numpy()
        for _ in range(3):                       # warm-up at full size: buffers allocated, kernel specialised
            regex.scan_host(host_np, off)
        e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        n_e = 0
        for _ in range(e_steps):
            n_e = len(regex.scan_host(host_np, off))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e_steps
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # what bounds it: the plain pinned H2D copy of the same buffer, timed alone
        d_tmp = torch.empty(e_size, dtype=torch.uint8, device="cuda")
        d_tmp.copy_(h_text, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_tmp.copy_(h_text, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = e_size / (time.perf_counter() - t0) / 1e9
        del d_tmp
        e2e = {"value": e_size * world / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": e_size, "d2h_bytes_per_step": 16 * n_e,
               "bytes_scanned_per_rank": e_size, "ms_per_step": dt * 1e3, "pinned_h2d_copy_alone_gbs": h2d_gbs,
               "call": "rrx_scan_host (C ABI) on a pinned host buffer: H2D copy, scan kernels, D2H of the match list into pinned memory"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    k_ms = float(np.mean(scan_kernel_ms)) if scan_kernel_ms else float(np.mean(kern_ms))
    alg_bytes = size + 16 * n_matches
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "scanned GB/s, bit-exact lazy matches", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": f"u32x{info['words']} bitsets", "data": "synthetic",
        "config": {"workload": f"{cfg.name}: regex {cfg.pattern!r} over {size} bytes per GPU ({cfg.note})",
                   "bytes_per_gpu": size, "matches_total": total_matches, "l2": "input larger than L2 (126 MB), no flush needed" if size > (256 << 20) else "input smaller than L2",
                   "states_raw": info["n_raw"], "states_flat": info["n_flat"], "states_device": info["n_device"], "words": info["words"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(cfg.name, size),
                     "kernel": "rrx_scan_jit" if is_jit else "rrx_scan_kernel", "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "step_ms_device": float(np.mean(kern_ms)),
                     "note": "kernel_ms: CUDA events around the scan kernel's launch alone; a step = 2 memsets + scan kernel + 3-kernel ordering pass"
                             + ("" if world == 1 else " + message copy + NCCL all-gather + D2H of the gathered messages; steps software-pipelined (step k+1 enqueued before the host collects step k)"),
                     "peak_source": peak_src},
        "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks,
        "kernel_variant": "run-time specialised (NVRTC)" if is_jit else "pre-compiled generic",
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, 32 << 20)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
