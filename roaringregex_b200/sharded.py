"""Sharded scans: one contiguous byte range per GPU.

Every rank scans its shard speculatively with an EMPTY entry set, then the ranks exchange one tiny
message each — the shard's composite transfer function (1+3*W words) — with ONE allgather, fold
the summaries of the ranks to their left into their exact entry set, and repair their result:

  * entry set empty, or dead within the shard's first bytes  -> nothing or a few matches to add;
  * otherwise                                                  -> the shard's PREFIX is rescanned
    with the exact entry until the exact and the speculative scan agree on the automaton state
    (from there on they agree forever); the exact prefix matches replace the speculative ones;
  * if they never agree inside the prefix window               -> the whole shard is rescanned.

Lazy start scans that run off a shard's left edge use `halo` bytes of left context resident on the
rank; scans that exhaust even that are handed to the rank on the left (second, rare exchange).

The collective is whatever `comm` provides (torch.distributed: NCCL on GPUs, gloo in the CPU
tests); the per-shard work is done by a `backend` — DeviceBackend (librrx, the product) or, in the
CPU test-suite only, an emulator-backed stand-in.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import Scanner, RRegex, lib, _check, SOFT_VERIFY_EDGE

ERR_VERIFY_EDGE, ERR_VERIFY_BUDGET = 16, 32      # rrx_tile.cuh

PREFIX_WINDOW = 1 << 20          # bytes of a shard rescanned with the exact entry before giving up
PENDING_RECORD = 8 + 8 + 4 * 32  # sizeof(PendingStart)


@dataclass
class ShardScan:
    """Outcome of one scan of (part of) a shard."""
    n: int                      # matches found
    matches: np.ndarray | None  # (n,2) int64 on the host, or None if left on the device
    summary: np.ndarray         # composite of the scanned range [kind, b, U, P]
    pending: list = field(default_factory=list)   # [(match_index, global_next_offset, R words)]
    handle: object = None       # backend-specific (device list)
    want_fallback: bool = False  # thread-owned scan (rrx_info.donly) not valid for this shard: everybody rescans without it


class DeviceBackend:
    """Scans a shard resident in HBM through librrx (the product path)."""

    def __init__(self, regex: RRegex, d_text_ptr: int, shard_len: int, base: int, halo: int, capacity: int, stream: int = 0):
        self.regex, self.ptr, self.len, self.base, self.halo, self.stream = regex, d_text_ptr, shard_len, base, halo, stream
        self.capacity = capacity
        self.scanner = Scanner(regex, shard_len, capacity)
        self.scanners = [self.scanner, None]      # the second one is created when steps are pipelined
        self._msgs = [None, None]
        self._turn = 0
        self.prefix_scanner = None
        self.words = regex.info["words"]

    def scan(self, length: int, entry, fetch: bool, prefix: bool = False, scanner=None, no_donly: bool = False) -> ShardScan:
        sc = scanner or self.scanner
        if prefix:
            if self.prefix_scanner is None:
                self.prefix_scanner = Scanner(self.regex, min(self.len, PREFIX_WINDOW), min(self.len, PREFIX_WINDOW) + 16, jit=False)
            sc = self.prefix_scanner
        sc.scan(self.ptr, length, self.base, self.halo, entry, self.stream, no_donly=no_donly)
        r = sc.result(self.stream)
        if r["error_bits"]:
            raise RuntimeError(f"scan kernel error bits {r['error_bits']}")
        if r["n_matches"] > r["capacity"]:
            raise RuntimeError("match buffer too small for this shard")
        pend = self._fetch_pending(sc, r["n_pending"])
        # the thread-owned scan assumed "nothing left of the readable text" (soft bit) although there is, or fell
        # back by itself (budget): its composite is not comparable with the other ranks' — all redo without it
        want = (not no_donly) and (bool(r["soft_bits"] & SOFT_VERIFY_EDGE and self.base - self.halo > 0) or sc.used_fallback)
        return ShardScan(n=r["n_matches"], matches=sc.matches(self.stream) if fetch else None,
                         summary=sc.shard_summary(self.stream), pending=pend, handle=sc, want_fallback=want)

    def _fetch_pending(self, sc, n_pending: int) -> list:
        pend = []
        if n_pending:
            buf = np.zeros(n_pending * PENDING_RECORD, dtype=np.uint8)
            n = C.c_uint64(0)
            _check(lib().rrx_scanner_copy_pending(sc._h, buf.ctypes.data, n_pending, C.byref(n), self.stream))
            if int(n.value) < n_pending:
                raise RuntimeError("too many start scans cross the shard's left edge (pending list overflow)")
            for i in range(int(n.value)):
                rec = buf[i * PENDING_RECORD:(i + 1) * PENDING_RECORD]
                mi, nxt = np.frombuffer(rec[:16].tobytes(), dtype=np.int64)
                pend.append((int(mi), int(nxt), np.frombuffer(rec[16:].tobytes(), dtype=np.uint32).copy()))   # offsets are global
        return pend

    def scan_gather_begin(self, comm):
        """Enqueue — without waiting for anything — the speculative scan of the whole shard, a device-side
        copy of [composite | counters], the all-gather of that message and the device-to-host copy of the
        gathered messages, all on one stream.  Two sets of buffers alternate, so the next step may be
        enqueued before this one has been collected (scan_gather_end).  Returns a token, or None if
        `comm` cannot gather device buffers."""
        if not hasattr(comm, "allgather_device_begin"):
            return None
        import torch
        slot = self._turn
        self._turn ^= 1
        if self.scanners[slot] is None:
            self.scanners[slot] = Scanner(self.regex, self.len, self.capacity)
        sc, W = self.scanners[slot], self.words
        nw = 1 + 3 * W + 8
        if self._msgs[slot] is None:
            self._msgs[slot] = torch.zeros(nw, dtype=torch.int32, device="cuda")
        sc.scan(self.ptr, self.len, self.base, self.halo, None, self.stream)
        sc.message_device(self._msgs[slot].data_ptr(), nw, self.stream)
        return dict(slot=slot, scanner=sc, gather=comm.allgather_device_begin(self._msgs[slot], slot))

    def scan_gather_end(self, comm, token, fetch: bool):
        """Wait for the step enqueued by scan_gather_begin: ONE synchronisation (on the event behind
        its device-to-host copy).  -> (ShardScan, (world, 1+3W+1) uint32 messages as scan_sharded exchanges)."""
        sc, W = token["scanner"], self.words
        g = comm.allgather_device_end(token["gather"])          # (world, nw) uint32 on the host
        mine = g[comm.dist.get_rank()]
        ctr = mine[1 + 3 * W:]
        err, n_matches, n_pending = int(ctr[1]), int(ctr[6]) | (int(ctr[7]) << 32), int(ctr[4]) | (int(ctr[5]) << 32)
        if err & ~(ERR_VERIFY_EDGE | ERR_VERIFY_BUDGET):
            raise RuntimeError(f"scan kernel error bits {err}")
        if n_matches > self.capacity:
            raise RuntimeError("match buffer too small for this shard")
        pend = self._fetch_pending(sc, n_pending)               # rare; no collective involved
        # every rank derives every rank's "redo without the thread-owned scan" flag from the gathered error words
        errs = g[:, 1 + 3 * W + 1].astype(np.int64)
        left_text = np.array([r > 0 for r in range(g.shape[0])])         # ranks with text left of their halo
        flags = ((errs & ERR_VERIFY_BUDGET) != 0) | (((errs & ERR_VERIFY_EDGE) != 0) & left_text)
        spec = ShardScan(n=n_matches, matches=sc.matches(self.stream) if fetch and not flags.any() else None,
                         summary=mine[:1 + 3 * W].copy(), pending=pend, handle=sc, want_fallback=bool(flags[comm.dist.get_rank()]))
        msgs = np.concatenate([g[:, :1 + 3 * W], np.minimum(g[:, 1 + 3 * W + 4:1 + 3 * W + 5], 0xffffffff),
                               flags[:, None].astype(np.uint32)], axis=1).astype(np.uint32)
        return spec, msgs

    def head(self, scan: ShardScan, k: int) -> np.ndarray:
        """First k matches of a scan left on the device."""
        k = min(k, scan.n)
        out = np.empty((k, 2), dtype=np.int64)
        _check(lib().rrx_scanner_copy_matches(scan.handle._h, out.ctypes.data, k, self.stream))
        return out

    def continue_starts(self, records):
        """records: [(global_next_offset, R words)] -> (starts list, updated records)."""
        n = len(records)
        buf = np.zeros(n * PENDING_RECORD, dtype=np.uint8)
        for i, (nxt, R) in enumerate(records):
            buf[i * PENDING_RECORD + 8:i * PENDING_RECORD + 16] = np.frombuffer(np.int64(nxt).tobytes(), dtype=np.uint8)
            buf[i * PENDING_RECORD + 16:(i + 1) * PENDING_RECORD] = np.frombuffer(np.asarray(R, dtype=np.uint32).tobytes(), dtype=np.uint8)
        starts = np.zeros(n, dtype=np.int64)
        _check(lib().rrx_continue_starts(self.regex._h, self.ptr, self.len, self.halo, self.base, buf.ctypes.data, n,
                                         starts.ctypes.data, self.stream))
        out = []
        for i in range(n):
            rec = buf[i * PENDING_RECORD:(i + 1) * PENDING_RECORD]
            out.append((int(np.frombuffer(rec[8:16].tobytes(), dtype=np.int64)[0]), np.frombuffer(rec[16:].tobytes(), dtype=np.uint32).copy()))
        return starts.tolist(), out


def fold_entry(regex: RRegex, summaries: np.ndarray, rank: int):
    """Exact entry set of `rank` from all ranks' summaries ((world, 1+3W) uint32).  -> (entry, exact)"""
    w = regex.info["words"]
    s = np.ascontiguousarray(summaries, dtype=np.uint32)
    entry = np.zeros(w, dtype=np.uint32)
    exact = C.c_int(0)
    _check(lib().rrx_fold_entry(regex._h, s.ctypes.data, s.shape[0], rank, entry.ctypes.data, C.byref(exact)))
    return entry, bool(exact.value)


def fold_entries(regex: RRegex, summaries: np.ndarray):
    """fold_entry for every rank in one call -> [(entry, exact)] * world."""
    w = regex.info["words"]
    s = np.ascontiguousarray(summaries, dtype=np.uint32)
    world = s.shape[0]
    entries = np.zeros((world, w), dtype=np.uint32)
    exact = np.zeros(world, dtype=np.int32)
    _check(lib().rrx_fold_entries(regex._h, s.ctypes.data, world, entries.ctypes.data, exact.ctypes.data))
    return [(entries[r], bool(exact[r])) for r in range(world)]


def apply_summary(regex: RRegex, summary: np.ndarray, entry: np.ndarray) -> np.ndarray:
    w = regex.info["words"]
    out = np.zeros(w, dtype=np.uint32)
    s = np.ascontiguousarray(summary, dtype=np.uint32)
    e = np.ascontiguousarray(entry, dtype=np.uint32)
    _check(lib().rrx_apply_summary(regex._h, s.ctypes.data, e.ctypes.data, out.ctypes.data))
    return out


@dataclass
class ShardResult:
    """Exact matches of one shard: `head` (host, repaired prefix) followed by the speculative scan's
    matches from index `skip` on (device list `tail`, or host array when fetched)."""
    count: int
    head: np.ndarray
    skip: int
    tail: ShardScan
    rescans: int = 0             # 0 none, 1 prefix only, 2 whole shard
    pending: list = field(default_factory=list)   # [(where, index, global_next_offset, R)], where in {"head","tail"}
    patches: list = field(default_factory=list)   # [(where, index, start)] starts found by ranks to the left

    def to_host(self, backend) -> np.ndarray:
        t = self.tail.matches if self.tail.matches is not None else backend.head(self.tail, self.tail.n)
        out = np.concatenate([self.head, t[self.skip:]]) if (len(self.head) or self.skip) else t.copy()
        for where, idx, s in self.patches:
            out[idx if where == "head" else len(self.head) + idx - self.skip, 0] = s
        return out


def scan_sharded_begin(regex: RRegex, backend, comm):
    """Enqueue one sharded step on the device and return at once (a token for scan_sharded_finish), or
    None when the backend/communicator pair has no asynchronous path.  Steps begun back to back
    overlap the host side of one with the device side of the next."""
    return backend.scan_gather_begin(comm) if hasattr(backend, "scan_gather_begin") else None


def scan_sharded_finish(regex: RRegex, backend, comm, rank: int, world: int, token, fetch: bool = False) -> "ShardResult":
    return scan_sharded(regex, backend, comm, rank, world, fetch, _token=token)


def scan_sharded(regex: RRegex, backend, comm, rank: int, world: int, fetch: bool = False, _token=None) -> ShardResult:
    """The per-rank procedure (see _scan_sharded).  A program with the thread-owned candidate scan (rrx_info.donly) runs
    it first; if any rank reports that it does not hold for its shard — a verification walk reached the left edge of
    the rank's readable text although there is text further left, or ran out of budget — every rank sees that in the
    gathered messages and all of them redo the step with the row-bounded scan (rare: needs a verification walk
    longer than the halo)."""
    res = _scan_sharded(regex, backend, comm, rank, world, fetch, _token, False)
    if res is None:
        res = _scan_sharded(regex, backend, comm, rank, world, fetch, None, True)
    return res


def _scan_sharded(regex: RRegex, backend, comm, rank: int, world: int, fetch: bool, _token, no_donly: bool):
    """The per-rank procedure.  `comm.allgather(np.ndarray) -> (world, ...) ndarray` is the only
    collective on the data path: ONE message of 1+3W+1 words per rank (the shard summary and the
    number of unfinished start scans).  Everything else each rank derives locally from the gathered
    summaries (every rank can fold every other rank's entry), so in the common case — no entry set
    survives into any shard, no start scan crosses a halo — a step has exactly one collective.
    `comm.allgather_obj(obj)` (pickled, tiny) serves the rare hand-over of unfinished start scans."""
    kw = {"no_donly": True} if no_donly else {}
    if _token is None and not no_donly:
        _token = scan_sharded_begin(regex, backend, comm)
    if _token is not None:
        spec, first_gather = backend.scan_gather_end(comm, _token, fetch)    # scan + exchange already enqueued, one sync
    else:
        spec, first_gather = backend.scan(backend.len, None, fetch, **kw), None     # speculative: empty entry
    advertised, result, last_entry = spec, None, None
    consumed = False              # a whole-shard rescan has overwritten the speculative scan's device buffers
    for _round in range(max(world, 1)):
        if _round == 0 and first_gather is not None:
            gathered = first_gather
        else:
            msg = np.concatenate([advertised.summary.astype(np.uint32),
                                  np.array([len(spec.pending), int(spec.want_fallback)], dtype=np.uint32)])
            gathered = comm.allgather(msg)
        summaries, spec_pending, fallback = gathered[:, :-2], gathered[:, -2], gathered[:, -1]
        if fallback.any() and not no_donly:
            return None                                           # same data on every rank: everybody redoes the step
        folds = fold_entries(regex, summaries)
        entry, _ = folds[rank]
        if result is None or not np.array_equal(entry, last_entry):
            last_entry = entry
            if entry.any():
                result = _repair(regex, backend, spec, entry, fetch, kw, force_full=consumed)
                if result.rescans == 2:
                    advertised = result.tail                      # now computed with the exact entry
                    consumed = True                               # later rounds must not splice onto the stale speculative list
            else:
                result = ShardResult(spec.n, np.zeros((0, 2), np.int64), 0, spec, 0,
                                     [("tail", i, nxt, R) for i, nxt, R in spec.pending])
        if all(ex for _e, ex in folds):
            break                                                 # nobody folded through an opaque summary
    any_repair = any(e.any() for e, _ex in folds)
    if any_repair or spec_pending.any():
        _resolve_pending(backend, comm, rank, world, result)
    return result


def _repair(regex, backend, spec: ShardScan, entry: np.ndarray, fetch: bool, kw: dict, force_full: bool = False) -> ShardResult:
    if force_full:
        kw2 = dict(kw)
        if isinstance(backend, DeviceBackend) and spec.handle is not None:
            kw2["scanner"] = spec.handle
        full = backend.scan(backend.len, entry, fetch, **kw2)
        return ShardResult(full.n, np.zeros((0, 2), np.int64), 0, full, 2, [("tail", i, nxt, R) for i, nxt, R in full.pending])
    window = min(backend.len, PREFIX_WINDOW)
    if window < backend.len:
        window -= window % 4096
    a = backend.scan(window, None, True, prefix=True, **kw)       # speculative state after the window
    b = backend.scan(window, entry, True, prefix=True, **kw)      # exact state after the window
    sa = apply_summary(regex, a.summary, np.zeros_like(entry))
    sb = apply_summary(regex, b.summary, entry)
    if np.array_equal(sa, sb) or window == backend.len:
        skip = a.n                                                # speculative matches ending inside the window
        pend = [("tail", i, nxt, R) for i, nxt, R in spec.pending if i >= skip] + [("head", i, nxt, R) for i, nxt, R in b.pending]
        return ShardResult(len(b.matches) + spec.n - skip, b.matches, skip, spec, 1, pend)
    # the entry's influence never died: rescan — into the buffers of the speculative scan it replaces
    # (with pipelined steps the other scanner may already hold the next step)
    kw2 = dict(kw)
    if isinstance(backend, DeviceBackend) and spec.handle is not None:
        kw2["scanner"] = spec.handle
    full = backend.scan(backend.len, entry, fetch, **kw2)
    return ShardResult(full.n, np.zeros((0, 2), np.int64), 0, full, 2, [("tail", i, nxt, R) for i, nxt, R in full.pending])


def _resolve_pending(backend, comm, rank, world, result: ShardResult):
    """Start scans that ran off the left edge of the rank's readable text travel left, one rank
    per hop, until some rank finds the start.  Usually there are none."""
    counts = comm.allgather(np.array([len(result.pending)], dtype=np.uint32))
    if not counts.any():
        return
    flight = [(rank, where, idx, nxt, R) for where, idx, nxt, R in result.pending]   # records I must hand left
    found = []
    for hop in range(1, world):
        everyone = comm.allgather_obj(flight)
        todo = [rec for recs in everyone for rec in recs if rec[0] - hop == rank]      # owner - hop == me
        flight = []
        if todo:
            starts, upd = backend.continue_starts([(nxt, R) for _o, _w, _i, nxt, R in todo])
            for (owner, where, idx, _n, _r), s, (nxt2, R2) in zip(todo, starts, upd):
                if s >= 0:
                    found.append((owner, where, idx, s))
                elif s == -1:
                    flight.append((owner, where, idx, nxt2, R2))
    fixes = comm.allgather_obj(found)
    result.patches = [(where, idx, s) for recs in fixes for (owner, where, idx, s) in recs if owner == rank]


class TorchComm:
    """torch.distributed as the collective provider (NCCL with device tensors, gloo with CPU ones)."""

    def __init__(self, dist, device):
        import torch
        self.dist, self.device = dist, torch.device(device)

    def allgather(self, arr: np.ndarray) -> np.ndarray:
        """One all_gather_into_tensor on preallocated buffers; pinned staging on both sides when the
        collective runs on a GPU (one H2D, one collective, one D2H per call)."""
        import torch
        a = np.ascontiguousarray(arr).view(np.int32)
        n, world = a.size, self.dist.get_world_size()
        key = (n, world)
        if getattr(self, "_key", None) != key:
            cuda = self.device.type == "cuda"
            self._key = key
            self._h_in = torch.empty(n, dtype=torch.int32, pin_memory=cuda)
            self._h_out = torch.empty(world * n, dtype=torch.int32, pin_memory=cuda)
            self._d_in = torch.empty(n, dtype=torch.int32, device=self.device)
            self._d_out = torch.empty(world * n, dtype=torch.int32, device=self.device)
        self._h_in.numpy()[:] = a
        self._d_in.copy_(self._h_in, non_blocking=True)
        self.dist.all_gather_into_tensor(self._d_out, self._d_in)
        self._h_out.copy_(self._d_out, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return self._h_out.numpy().view(np.uint32).reshape(world, n).copy()

    def allgather_device_begin(self, d_msg, slot: int = 0):
        """All-gather of a device tensor + copy of the result to pinned host memory, both enqueued on the
        current stream; an event marks the end.  Two buffer sets (slot 0/1)."""
        import torch
        if self.device.type != "cuda":
            raise RuntimeError("device all-gather needs a CUDA communicator")
        n, world = d_msg.numel(), self.dist.get_world_size()
        bufs = getattr(self, "_dbufs", None)
        if bufs is None or bufs["key"] != (n, world):
            bufs = self._dbufs = dict(key=(n, world),
                                      g_out=[torch.empty(world * n, dtype=d_msg.dtype, device=self.device) for _ in range(2)],
                                      g_host=[torch.empty(world * n, dtype=d_msg.dtype, pin_memory=True) for _ in range(2)],
                                      ev=[torch.cuda.Event() for _ in range(2)])
        self.dist.all_gather_into_tensor(bufs["g_out"][slot], d_msg)
        bufs["g_host"][slot].copy_(bufs["g_out"][slot], non_blocking=True)
        bufs["ev"][slot].record(torch.cuda.current_stream(self.device))
        return (slot, n, world)

    def allgather_device_end(self, handle) -> np.ndarray:
        slot, n, world = handle
        self._dbufs["ev"][slot].synchronize()
        return self._dbufs["g_host"][slot].numpy().view(np.uint32).reshape(world, n).copy()

    def allgather_device(self, d_msg) -> np.ndarray:
        return self.allgather_device_end(self.allgather_device_begin(d_msg, 0))

    def allgather_obj(self, obj):
        out = [None] * self.dist.get_world_size()
        self.dist.all_gather_object(out, obj)
        return out


class LocalComm:
    """Single-rank stand-in."""

    def allgather(self, arr):
        return np.asarray(arr)[None]

    def allgather_obj(self, obj):
        return [obj]
