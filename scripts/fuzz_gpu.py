#!/usr/bin/env python
"""Time-bounded fuzz of the CUDA path on the GPU box: random regexes x random texts through the C ABI (librrx.so) against
the oracle — `rrx_scan_host` (pre-compiled kernels), a scanner with the run-time specialised kernel (lean and, forced by
RRX_JIT_VARIANT=1 in a second process, straddle), and two-shard scans chained through the shard summary with the
documented redo when the thread-owned scan asks for it.  TEST INFRASTRUCTURE: the oracle is the checker here.

    python scripts/fuzz_gpu.py --seed 1 --minutes 3 [--jit-every 12]
    python scripts/fuzz_gpu.py --seed 1 --minutes 3 --mode extra      # anchors ^ $ and the non-overlapping view (one-shot calls)
    RRX_HOST_CHUNK=1048576 python scripts/fuzz_gpu.py --seed 1 --minutes 3 --mode stream   # streaming host ingestion, 3-6 MiB texts
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import roaringregex_b200 as rrx  # noqa: E402
from fuzz_emu import ALPHABETS, regex  # noqa: E402
from oracle import api as oracle_api  # noqa: E402


def two_shards(torch, r, sc, d, text, n, cut, ctx, want, stats):
    """left = [0, cut), right = [cut, n) with `ctx` bytes of left context and the left shard's exit as entry"""
    nw = r.info["words"]
    no_donly = False
    for attempt in range(2):
        sc.scan(d.data_ptr(), cut, no_donly=no_donly)
        lres = sc.result()
        left = sc.matches()
        comp = sc.shard_summary()
        lfall = sc.used_fallback
        if comp[0] & 0xff not in (0, 1):
            stats["open_summary"] += 1
            return None
        sc.scan(d.data_ptr() + cut, n - cut, base_offset=cut, left_context=ctx, entry=comp[1:1 + nw].copy(), no_donly=no_donly)
        rres = sc.result()
        if lres["error_bits"] or rres["error_bits"]:
            if rres["error_bits"] == 2 and not lres["error_bits"]:
                stats["pending_overflow"] += 1
                return None
            return ("shard-err", lres["error_bits"], rres["error_bits"])
        redo = (not no_donly) and (lfall or sc.used_fallback or (rres["soft_bits"] and ctx < cut))
        if not redo:
            break
        stats["shard_fallback"] += 1
        no_donly = True
    got = np.concatenate([left, sc.matches()])
    stats["shards"] += 1
    if len(got) != len(want) or not np.array_equal(got[:, 1], want[:, 1]):
        return ("shard-ends", len(got), len(want))
    if rres["n_pending"] == 0 and not np.array_equal(got, want):
        return ("shard-starts",)
    return None


def np_text(rng, letters, wts, n):
    p = np.asarray(wts, dtype=np.float64)
    return rng.choice(np.frombuffer(letters, dtype=np.uint8), size=n, p=p / p.sum()).astype(np.uint8)


def other_modes(a, torch):
    """extra: anchored patterns and RRX_MODE_NO_OVERLAP through rrx_scan_host / rrx_scan_device, thread-owned and
    row-bounded programs.  stream: rrx_scan_host on texts of several chunks (RRX_HOST_CHUNK), pageable and pinned."""
    rnd = random.Random(a.seed)
    rng = np.random.default_rng(a.seed)
    stats = dict(cases=0, anchored=0, no_overlap=0, stream=0, failures=0)
    t_end = time.time() + 60 * a.minutes
    it = 0
    while time.time() < t_end and stats["failures"] < 20:
        it += 1
        letters = rnd.choice(ALPHABETS)
        pat = regex(rnd, letters)
        ls = [chr(c) for c in letters if chr(c).isalnum()]
        anchored = a.mode == "extra" and rnd.random() < 0.5
        if anchored:
            pat = rnd.choice(["^" + pat, pat + "$", "^" + pat + "$", "(^|" + rnd.choice(ls) + ")" + pat, pat + "($|" + rnd.choice(ls) + ")"])
        try:
            o = oracle_api.Oracle(pat)
        except Exception:
            continue
        if o.n_raw > 500:
            continue
        progs = []
        try:
            for flags in (0, rrx.COMPILE_NO_DONLY):
                progs.append(rrx.RRegex(pat, flags))
        except rrx.RrxError as e:
            stats["failures"] += 1
            print("FAIL", "seed", a.seed, "iter", it, ("compile", pat, str(e)), flush=True)
            continue
        for _ in range(3):
            wts = [rnd.choice([1, 1, 5, 40, 200]) for _ in letters]
            if a.mode == "stream":
                n = rnd.choice([3 << 20, (3 << 20) + 1, (4 << 20) + 4097, (5 << 20) - 1, 6 << 20])
            else:
                n = rnd.choice([0, 1, 2, 127, 128, 129, 700, 4095, 4096, 4097, 9000, 33000, 70000])
            text = np_text(rng, letters, wts, n)
            tb = text.tobytes()
            want = o.lazy_matches_anchored(tb) if anchored else o.lazy_matches(tb)
            if len(want) > (1 << 21):
                continue
            stats["cases"] += 1
            bad = None
            for k, r in enumerate(progs):
                try:
                    got = r.scan_host(text)
                except rrx.RrxError as e:
                    bad = ("scan_host error", k, str(e))
                    break
                if got.shape != want.shape or not np.array_equal(got, want):
                    bad = ("scan_host", k, len(got), len(want))
                    break
                if a.mode == "stream":
                    stats["stream"] += 1
                    pinned = torch.from_numpy(text.copy()).pin_memory()
                    got = r.scan_host(pinned.numpy())
                    if got.shape != want.shape or not np.array_equal(got, want):
                        bad = ("scan_host pinned", k, len(got), len(want))
                        break
                    continue
                stats["anchored"] += int(anchored)
                if n:
                    d = torch.from_numpy(text.copy()).cuda()
                    got = r.scan_device(d.data_ptr(), n)
                    if got.shape != want.shape or not np.array_equal(got, want):
                        bad = ("scan_device", k, len(got), len(want))
                        break
                if not anchored:
                    stats["no_overlap"] += 1
                    wn = o.no_overlap_matches(tb)
                    got = r.scan_host(text, 0, rrx.MODE_NO_OVERLAP)
                    if got.shape != wn.shape or not np.array_equal(got, wn):
                        bad = ("no_overlap", k, len(got), len(wn))
                        break
            if bad:
                stats["failures"] += 1
                print("FAIL", "seed", a.seed, "iter", it, (pat, letters, wts, n) + tuple(bad), flush=True)
                np.save(os.path.join(ROOT, "gpurun_out", f"fuzz_gpu_fail_{a.mode}_{a.seed}_{it}.npy"), text)
                break
    print("fuzz_gpu", a.mode, "seed", a.seed, "chunk_env", os.environ.get("RRX_HOST_CHUNK", ""), stats, flush=True)
    return 1 if stats["failures"] else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=3.0)
    ap.add_argument("--jit-every", type=int, default=12, help="every k-th pattern also runs on a run-time specialised scanner")
    ap.add_argument("--mode", default="base", choices=["base", "extra", "stream"])
    a = ap.parse_args()
    import torch
    assert torch.cuda.is_available(), "needs a CUDA device"
    if a.mode != "base":
        return other_modes(a, torch)
    rnd = random.Random(a.seed)
    stats = dict(cases=0, jit=0, shards=0, shard_fallback=0, pending_overflow=0, open_summary=0, failures=0)
    t_end = time.time() + 60 * a.minutes
    NMAX = 70000
    d = torch.zeros(NMAX + 4096, dtype=torch.uint8, device="cuda")
    it = 0
    while time.time() < t_end and stats["failures"] < 20:
        it += 1
        letters = rnd.choice(ALPHABETS)
        pat = regex(rnd, letters)
        try:
            o = oracle_api.Oracle(pat)
        except Exception:
            continue
        if o.n_raw > 700:
            continue
        try:
            r = rrx.RRegex(pat)
        except rrx.RrxError as e:
            stats["failures"] += 1
            print("FAIL", "seed", a.seed, "iter", it, ("compile", pat, str(e)), flush=True)
            continue
        use_jit = it % a.jit_every == 0
        scanners = [rrx.Scanner(r, NMAX, NMAX + 64, jit=False)]
        if use_jit:
            scanners.append(rrx.Scanner(r, NMAX, NMAX + 64, jit=True))
            stats["jit"] += 1
        for _ in range(3):
            wts = [rnd.choice([1, 1, 5, 40, 200]) for _ in letters]
            n = rnd.choice([0, 1, 127, 128, 129, 700, 4095, 4096, 4097, 9000, 33000, 70000])
            text = np.frombuffer(bytes(rnd.choices(letters, weights=wts, k=n)), dtype=np.uint8)
            want = o.lazy_matches(text)
            if len(want) > NMAX:
                continue
            stats["cases"] += 1
            bad = None
            got = r.scan_host(text)
            if got.shape != want.shape or not np.array_equal(got, want):
                bad = ("scan_host", len(got), len(want))
            d[:n] = torch.from_numpy(text.copy()).cuda()
            d[n:n + 4096] = 0
            for sc in scanners:
                if bad:
                    break
                sc.scan(d.data_ptr(), n)
                res = sc.result()
                if res["error_bits"]:
                    bad = ("scanner-err", sc.is_jit, res["error_bits"])
                    break
                got = sc.matches()
                if got.shape != want.shape or not np.array_equal(got, want):
                    bad = ("scanner", sc.is_jit, sc.kernel_variant, len(got), len(want))
                    break
                if n >= 700:
                    cut = rnd.choice([(n // 2) & ~127, ((n // 2) & ~4095) or 128, 128, (n - 1) & ~15, ((n // 3) & ~15) or 16])
                    cut = min(max(cut, 16), (n - 1) & ~15)
                    ctx = rnd.choice([0, min(cut, 128), cut])
                    bad = two_shards(torch, r, sc, d, text, n, cut, ctx, want, stats)
                    if bad:
                        bad = bad + (sc.is_jit, sc.kernel_variant, cut, ctx)
            if bad:
                stats["failures"] += 1
                print("FAIL", "seed", a.seed, "iter", it, (pat, letters, wts, n) + tuple(bad), flush=True)
                np.save(os.path.join(ROOT, "gpurun_out", f"fuzz_gpu_fail_{a.seed}_{it}.npy"), text)
                break
    print("fuzz_gpu seed", a.seed, "variant_env", os.environ.get("RRX_JIT_VARIANT", ""), stats, flush=True)
    return 1 if stats["failures"] else 0


if __name__ == "__main__":
    sys.exit(main())
