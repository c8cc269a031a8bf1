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
    ap.add_argument("--config", default="2", help="1..5 or 2v (BASELINE.json configs; 2v = config 2 on variable-length lines)")
    ap.add_argument("--sub", default=None, help="comma list of configs measured as sub-records at their named sizes "
                    "(default at N=1 with --config 2: 1,2v,3,4,5; at N>1: 5 = config 5 strong); 'none' to skip")
    ap.add_argument("--prefix", type=int, default=256 << 20, help="bytes of each corpus the oracle re-scans for the parity digests")
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


def cpu_baseline(cfg, sample_per_thread: int, text=None, repeats: int = 1):
    """The oracle port (reference algorithm: per-active-state row OR, injection, flag test, backward
    start scan) on all host cores, each core scanning its own contiguous piece of the corpus."""
    from oracle import api as oracle_api
    oracle_api.build(ref=False)
    cores = os.cpu_count() or 1
    o = oracle_api.Oracle(cfg.pattern)
    n = sample_per_thread * cores
    if text is None:
        text = oracle_api.corpus(cfg, 0, n)
    o.time_parallel(text, min(sample_per_thread, 1 << 20), cores)         # warm
    sec, m = min(o.time_parallel(text, sample_per_thread, cores) for _ in range(max(repeats, 1)))
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
    # the same bytes as the GPU arm's step: the config's size (1 GiB for config 2, capped at 1 GiB for the
    # larger ones so that the run ends within minutes), split evenly over the host threads
    from oracle import api as oracle_api
    oracle_api.build(ref=False)
    cores = os.cpu_count() or 1
    total = args.size or min(cfg.length, 1 << 30)
    per_thread = max(total // cores, 1 << 16)
    text = oracle_api.corpus(cfg, 0, per_thread * cores)
    vals, res = [], None
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(cfg, per_thread, text=text)
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
            "config": {"workload": f"{cfg.name}: regex {cfg.pattern!r} over {nbytes} bytes per step (the first {nbytes >> 20} MiB of the corpus, "
                                   f"{per_thread >> 20} MiB per host thread), full LAZY-v1 match lists", "bytes_per_step": nbytes},
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
# Parity at scale: the device digest of the ordered match list against (a) the oracle's
# digest over a prefix of the same corpus, (b) a count that does not come from any scan kernel.
def closed_form_count(torch, cfg, d_text, size, off=0):
    """Number of LAZY-v1 matches the corpus must contain, from plain torch reductions over the device text
    (iid corpora: the regex's last byte preceded by what the regex needs) or from the generator's own
    record kinds (log corpora).  Independent of librrx.  Shard-local for off > 0 (config 5 only counts
    q's; the caller subtracts those before the corpus's first z)."""
    from oracle import api as oracle_api
    name = cfg.name
    if name == "config2":
        assert off == 0 and size % 128 == 0
        return oracle_api.records_valid(cfg.seed, size // 128)
    if name == "config2v":
        assert off == 0 and size % 2048 == 0
        return oracle_api.lines_valid(cfg.seed, size // 2048)
    a, b, c, e, q, z = (ord(ch) for ch in "abceqz")
    total, first_z, CH = 0, None, 1 << 30
    for lo in range(0, size, CH):
        hi = min(size, lo + CH)
        t = d_text[lo:hi]
        if name == "config1":
            total += int((t == e).sum())
        elif name in ("config3", "config4"):
            last = b if name == "config3" else c
            p = d_text[max(lo - 1, 0):hi - 1] if lo else torch.cat([torch.zeros(1, dtype=torch.uint8, device=t.device), d_text[:hi - 1]])
            ok = (p == a) if name == "config3" else ((p == a) | (p == b))
            total += int(((t == last) & ok).sum())
        elif name == "config5":
            isq = t == q
            if first_z is None:
                zpos = torch.nonzero(t == z)
                if len(zpos):
                    first_z = lo + int(zpos[0])
                    total += int(isq[first_z - lo:].sum())
            else:
                total += int(isq.sum())
        else:
            return None
    return (total, first_z) if name == "config5" else total


def oracle_prefix_digest(cfg, prefix):
    """Digest of the oracle's LAZY-v1 list over corpus[0:prefix).  Configs 3/4 (bounded match length) are
    scanned in parallel pieces with 4 KiB of overlap, which is exact for them; the others in one piece."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import api as oracle_api
    o = oracle_api.Oracle(cfg.pattern)
    t0 = time.perf_counter()
    if cfg.name in ("config3", "config4"):
        cores = os.cpu_count() or 1
        piece = -(-prefix // cores)
        def work(i):
            lo, hi = i * piece, min(prefix, (i + 1) * piece)
            if lo >= hi:
                return np.zeros((0, 2), np.int64)
            ov = min(lo, 4096)
            m = o.lazy_matches(oracle_api.corpus(cfg, lo - ov, hi - lo + ov), base=lo - ov)
            return m[m[:, 1] > lo]
        with ThreadPoolExecutor(cores) as ex:
            m = np.concatenate(list(ex.map(work, range(cores))))
    else:
        m = o.lazy_matches(oracle_api.corpus(cfg, 0, prefix))
    return oracle_api.digest(m), time.perf_counter() - t0


def parity_record(rrx, torch, cfg, sc, d_text, size, off, n_matches, prefix, sh, closed=True):
    prefix = min(prefix, size)
    full = sc.digest(1 << 62, sh)
    pre = sc.digest(off + prefix, sh) if off == 0 else None
    rec = {"digest": "count, sum and xor (mod 2^64) of sm64(sm64(start)+end) over the ordered list, computed on the device (rrx_hash.cu)",
           "hash_gpu": [full[0], f"{full[1]:016x}", f"{full[2]:016x}"], "order_violations": full[3], "count_gpu": n_matches}
    ok = full[0] == n_matches and full[3] == 0
    if pre is not None:
        (oc, os_, ox, ob), secs = oracle_prefix_digest(cfg, prefix)
        rec.update({"prefix_bytes": prefix, "hash_gpu_prefix": [pre[0], f"{pre[1]:016x}", f"{pre[2]:016x}"],
                    "hash_oracle_prefix": [oc, f"{os_:016x}", f"{ox:016x}"], "oracle_seconds": round(secs, 2)})
        ok = ok and (pre[0], pre[1], pre[2]) == (oc, os_, ox) and ob == 0
    if closed:
        cf = closed_form_count(torch, cfg, d_text, size, off)
        if isinstance(cf, tuple):
            cf = cf[0]
        rec["count_closed_form"] = cf
        ok = ok and (cf is None or cf == n_matches)
    rec["ok"] = bool(ok)
    return rec


def _scan_setup(rrx, torch, cfg, size, off, halo, sh):
    d_buf = torch.empty(halo + size + 256, dtype=torch.uint8, device="cuda")
    d_text = d_buf[halo:]
    rrx.corpus_device(cfg, d_buf.data_ptr(), off - halo, halo + size, sh)
    torch.cuda.synchronize()
    return d_buf, d_text


def sub_record(rrx, torch, cfg, peak, prefix, sh, steps=3):
    """One configuration at its named size on one GPU: throughput, the scan kernel's own time, roofline
    fraction, match count and the parity digests."""
    size = cfg.length
    regex = rrx.RRegex(cfg.pattern)
    d_buf, d_text = _scan_setup(rrx, torch, cfg, size, 0, 0, sh)
    # match capacity as rrx_scan_device sizes it (one pair per 64 bytes): half of it is handed out as fixed staging slots,
    # 32 per 4 KiB tile, so that tiles with the corpora's 15-16 matches need no allocation
    cap = max(1 << 16, size // 64)
    sc = rrx.Scanner(regex, size, cap)
    sc.scan(d_text.data_ptr(), size, 0, 0, None, sh)
    res = sc.result(sh)
    if res["n_matches"] > cap:
        cap = res["n_matches"] + 1024
        del sc
        torch.cuda.empty_cache()
        sc = rrx.Scanner(regex, size, cap)
    sc.enable_timing(True)
    for _ in range(3):          # (the scanner picks its kernel variant by what the previous scan saw: let it settle)
        sc.scan(d_text.data_ptr(), size, 0, 0, None, sh)
        sc.result(sh)
    torch.cuda.synchronize()
    step_ms, kern_ms = [], []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        sc.scan(d_text.data_ptr(), size, 0, 0, None, sh)
        b.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        step_ms.append(a.elapsed_time(b))
        kern_ms.append(sc.kernel_ms())
    res = sc.result(sh)
    assert res["error_bits"] == 0 and res["n_matches"] <= cap, res
    n = res["n_matches"]
    k_ms, s_ms = float(np.mean(kern_ms)), float(np.mean(step_ms))
    alg = size + 16 * n
    info = regex.info
    rec = {"workload": f"{cfg.name}: regex {cfg.pattern!r} over {size} bytes ({cfg.note})", "bytes": size, "matches": n,
           "value": size / (s_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": s_ms, "kernel_ms": k_ms, "steps": steps,
           "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "frac": alg / (k_ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": alg},
           "kernel_variant": ("run-time specialised (NVRTC), " + ("straddle" if sc.kernel_variant else "lean") + " kernel") if sc.is_jit else "pre-compiled generic",
           "states_device": info["n_device"], "words": info["words"],
           "parity": parity_record(rrx, torch, cfg, sc, d_text, size, 0, n, prefix, sh)}
    if cfg.name in ("config3", "config4"):
        rec["parity"]["note"] = "the compiled reference cannot run this regex (>256 raw states, SURVEY.md §0.4): the oracle is the restated scanner, itself checked on scaled shapes and against Python re"
    return rec


def make_nccl_comm(rrx, torch, dist, rank, world):
    """A raw ncclComm_t for rrx_sharded_* (the unique id travels over torch.distributed)."""
    cached = getattr(make_nccl_comm, "_c", None)
    if cached is None:
        def bcast(raw):
            t = torch.tensor(list(raw), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        cached = make_nccl_comm._c = rrx.NcclComm(rank, world, bcast)
    return cached


def sub_record_sharded(rrx, torch, dist, cfg, peak, prefix, sh, rank, world, local, steps=3):
    """Config 5 as BASELINE.json names it: ONE 64 GiB corpus cut into `world` contiguous shards (strong scaling), scanned
    through the C ABI's sharded scan (rrx_sharded_*).  Sum/xor digests add up over the shards, so the whole-job digest
    must equal the 1-GPU one."""
    size = (cfg.length // world) & ~4095
    off = rank * size
    halo = min(off, 1 << 20)
    regex = rrx.RRegex(cfg.pattern)
    d_buf, d_text = _scan_setup(rrx, torch, cfg, size, off, halo, sh)
    cap = max(1 << 16, size // 64)
    probe = rrx.Scanner(regex, size, cap)
    probe.scan(d_text.data_ptr(), size, off, halo, None, sh)
    cap = max(cap, probe.result(sh)["n_matches"] + (1 << 16))
    del probe
    torch.cuda.empty_cache()
    ss = rrx.ShardedScanner(regex, size, cap, make_nccl_comm(rrx, torch, dist, rank, world))
    ss.enable_timing(True)
    res = None
    for _ in range(3):
        res = ss.scan(d_text.data_ptr(), size, off, halo, sh)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):          # pipelined: begin k+1 before finish k
        ss.begin(d_text.data_ptr(), size, off, halo, sh)
        if k:
            res = ss.finish()
    res = ss.finish()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    k_ms = ss.kernel_ms()
    dg = ss.digest(res, 1 << 62, sh)
    cf = closed_form_count(torch, cfg, d_text, size, off)
    t = torch.tensor([dg[0], dg[1] >> 32, dg[1] & 0xffffffff, dg[2] >> 32, dg[2] & 0xffffffff, dg[3], cf[0], -1 if cf[1] is None else cf[1],
                      res["rescans"], res["patched_starts"], res["n_matches"]], device="cuda", dtype=torch.int64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    tm = torch.tensor([ms, k_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, k_ms = float(tm[0].item()), float(tm[1].item())
    g = [v.tolist() for v in g]
    total = sum(v[0] for v in g)
    ssum = sum((v[1] << 32) | v[2] for v in g) % (1 << 64)
    sx = 0
    for v in g:
        sx ^= (v[3] << 32) | v[4]
    # closed form: q's after the corpus's first z.  Shard r counted the q's after ITS first z; the q's of shards right of
    # the globally first z that precede their own first z are owed too.
    cf_total = sum(v[6] for v in g)
    zs = [v[7] for v in g]
    if any(z >= 0 for z in zs):
        first_rank = next(i for i, z in enumerate(zs) if z >= 0)
        lo_fix = torch.tensor([0], device="cuda", dtype=torch.int64)
        if rank > first_rank:
            zloc = zs[rank] if zs[rank] >= 0 else size          # (closed_form_count reports shard-local offsets)
            lo_fix[0] = int((d_text[:zloc] == ord("q")).sum())
        dist.all_reduce(lo_fix)
        cf_total += int(lo_fix.item())
    full_bytes = size * world
    rec = {"workload": f"{cfg.name}: regex {cfg.pattern!r}, ONE corpus of {full_bytes} bytes cut into {world} contiguous shards (strong scaling), "
                       "scanned through rrx_sharded_* (C ABI; NCCL all-gather + fold on the device)",
           "bytes_total": full_bytes, "bytes_per_gpu": size, "matches": total, "value": full_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
           "ms_per_step": ms, "kernel_ms": k_ms, "steps": steps, "scaling": "strong",
           "roofline_frac_per_gpu": (size + 16 * total / world) / (k_ms * 1e-3) / 1e9 / peak,
           "rescans_per_rank": [v[8] for v in g], "start_patches_per_rank": [v[9] for v in g],
           "parity": {"hash_gpu": [total, f"{ssum:016x}", f"{sx:016x}"], "order_violations": sum(v[5] for v in g),
                      "count_closed_form": cf_total, "ok": bool(total == cf_total and sum(v[5] for v in g) == 0 and total == sum(v[10] for v in g)),
                      "note": "sum/xor digests are order-free: this triple must equal config5's hash_gpu in the 1-GPU record"}}
    del ss, d_text, d_buf
    return rec


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from roaringregex_b200.configs import CONFIGS
    key = args.config if args.config in CONFIGS else int(args.config)
    cfg = CONFIGS[key]
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
    strong = key == 5 and not args.size
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
    rrx.corpus_device(cfg, d_buf.data_ptr(), off - halo, halo + size, sh)
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
        # the C ABI's sharded scan (rrx_sharded_*, csrc/rrx_dist.cu): NCCL all-gather + fold on the device
        ncomm = make_nccl_comm(rrx, torch, dist, rank, world)
        probe = rrx.Scanner(regex, size, cap)
        probe.scan(d_text.data_ptr(), size, off, halo, None, sh)
        cap = max(cap, probe.result(sh)["n_matches"] + (1 << 16))
        is_jit = probe.is_jit
        del probe
        be = rrx.ShardedScanner(regex, size, cap, ncomm)
        last = {}

        def step():
            last["res"] = be.scan(d_text.data_ptr(), size, off, halo, sh)
        for _ in range(max(args.warmup, 4)):
            step()
        n_matches = last["res"]["n_matches"]
        launches_per_step = 4 * (1 + last["res"]["rescans"]) + 1         # + the fold kernel

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
        be.enable_timing(True)
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
        for k in range(args.steps):
            be.begin(d_text.data_ptr(), size, off, halo, sh)
            if k:
                last["res"] = be.finish()
                scan_kernel_ms.append(be.kernel_ms())
        last["res"] = be.finish()
        scan_kernel_ms.append(be.kernel_ms())
        assert last["res"]["n_matches"] == n_matches
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
               "call": "rrx_scan_host (C ABI) on a pinned host buffer: streaming ingestion — H2D in 64 MiB chunks on a copy stream, every chunk "
                       "scanned as it arrives (state carried), its matches copied back to pinned memory"}
        if world == 1:
            # the same call on PAGEABLE memory — what the C++ drop-in's getline buffer is (src/test/main.cpp:17-26)
            pg = np.array(host_np, copy=True)
            regex.scan_host(pg, off)
            t0 = time.perf_counter()
            for _ in range(2):
                n_p = len(regex.scan_host(pg, off))
            dtp = (time.perf_counter() - t0) / 2
            assert n_p == n_e
            e2e["pageable"] = {"value": e_size / dtp / 1e9, "unit": "GB/s", "ms_per_step": dtp * 1e3,
                               "note": "same call, host buffer NOT page-locked (plain malloc'ed memory)"}
            del pg
        del h_text, host_np

    peak, peak_src = measured_peaks()
    # ---- parity of the list that was just timed: device digest vs oracle on a prefix, closed-form count ----
    parity = None
    if world == 1:
        parity = parity_record(rrx, torch, cfg, sc, d_text, size, off, n_matches, args.prefix, sh)
    # ---- the other configurations at their named sizes (sub-records) ------------------------------------
    sub_keys = []
    if args.sub is None:
        sub_keys = ([1, "2v", 3, 4, 5] if world == 1 else [5]) if key == 2 and not args.size else []
    elif args.sub != "none":
        sub_keys = [k if k in CONFIGS else int(k) for k in args.sub.split(",") if k]
    # parity at N>1: the ranks' digests add / xor up to the digest of the whole list; count vs the closed form
    if world > 1:
        dg = be.digest(last["res"], 1 << 62, sh)
        cf = closed_form_count(torch, cfg, d_text, size, 0) if off == 0 or cfg.mode == "iid" else None
        if cfg.name == "config2":          # records are position-independent in kind: count this rank's records
            from oracle import api as oracle_api
            cf = oracle_api.records_valid(cfg.seed, (off + size) // 128) - oracle_api.records_valid(cfg.seed, off // 128)
        t = torch.tensor([dg[0], dg[1] >> 32, dg[1] & 0xffffffff, dg[2] >> 32, dg[2] & 0xffffffff, dg[3], cf if isinstance(cf, int) else -1,
                          last["res"]["rescans"]], device="cuda", dtype=torch.int64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        g = [v.tolist() for v in g]
        tot, ssum, sx = sum(v[0] for v in g), sum((v[1] << 32) | v[2] for v in g) % (1 << 64), 0
        for v in g:
            sx ^= (v[3] << 32) | v[4]
        cfs = [v[6] for v in g]
        parity = {"hash_gpu": [tot, f"{ssum:016x}", f"{sx:016x}"], "order_violations": sum(v[5] for v in g),
                  "count_closed_form": sum(cfs) if all(c >= 0 for c in cfs) else None, "rescans_per_rank": [v[7] for v in g],
                  "note": "weak scaling: rank r scans bytes [r*size, (r+1)*size) of ONE corpus; sum/xor digests over the ranks = digest of the whole list"}
        parity["ok"] = bool(tot == total_matches and parity["order_violations"] == 0 and (parity["count_closed_form"] in (None, tot)))
    sc = timed_scanner = step = be = None        # release the headline's device buffers
    if world > 1:
        last.clear()
    d_text = d_buf = None
    torch.cuda.empty_cache()
    subs = {}
    for k in sub_keys:
        try:
            if world == 1:
                subs[CONFIGS[k].name] = sub_record(rrx, torch, CONFIGS[k], peak, args.prefix, sh)
            else:
                subs[CONFIGS[k].name + "_strong"] = sub_record_sharded(rrx, torch, dist, CONFIGS[k], peak, args.prefix, sh, rank, world, local)
        except Exception as e:      # a sub-record must not take the headline line down with it
            subs[CONFIGS[k].name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
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
                             + ("" if world == 1 else " + message copy + ncclAllGather + fold kernel + 160-byte status copy, through rrx_sharded_begin/finish (C ABI); steps pipelined (begin k+1 before finish k)"),
                     "peak_source": peak_src},
        "e2e": e2e, "gpu_launches": args.steps * launches_per_step, "clocks": clocks,
        "kernel_variant": "run-time specialised (NVRTC)" if is_jit else "pre-compiled generic",
        "parity": parity, "configs": subs,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, 32 << 20)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
