"""Sharded scans (roaringregex_b200/sharded.py) on the CPU: the per-rank procedure — speculative scan,
ONE allgather of shard summaries, fold, prefix repair / rescan, hand-over of start scans — driven
by the kernel emulator as per-shard backend.  World sizes 2..4 over threads, and world size 2 over
real processes with torch.distributed's gloo backend."""
import os
import random
import threading

import numpy as np
import pytest

import emu_api
from conftest import cpu_corpus
from roaringregex_b200 import sharded
from roaringregex_b200.configs import CONFIGS
from tests_support import apply_follow, decode_blob


class EmuBackend:
    """TEST INFRASTRUCTURE: per-shard scans by the host-side kernel emulator."""
    DONLY = True          # let the compiler choose the thread-owned candidate scan (the `mode` fixture flips it)

    def __init__(self, rrx, pattern, text, lo, hi, halo):
        self.regex = rrx.RRegex(pattern, 0 if self.DONLY else rrx.COMPILE_NO_DONLY)
        self.pattern, self.text = pattern, text
        self.base, self.len = lo, hi - lo
        self.halo = min(halo, lo)
        self.prog = decode_blob(self.regex.blob())

    def scan(self, length, entry, fetch, prefix=False, no_donly=False):
        buf = self.text[self.base - self.halo:self.base + length]
        donly = self.DONLY and not no_donly
        r = emu_api.scan(self.pattern, buf, left_ctx=self.halo, base=self.base, entry=entry, donly=donly)
        want_fb = donly and bool((r["err"] & 32) or (r["soft"] and self.base - self.halo > 0))
        assert r["err"] & ~32 == 0
        W = r["info"]["NW"]
        pend = []
        for i in range(r["pending"]):
            rec = r["pend_raw"][i * r["pend_size"]:(i + 1) * r["pend_size"]]
            mi, nxt = np.frombuffer(rec[:16].tobytes(), dtype=np.int64)
            pend.append((int(mi), self.base + int(nxt), np.frombuffer(rec[16:].tobytes(), dtype=np.uint32).copy()))
        return sharded.ShardScan(n=r["n"], matches=r["matches"], summary=r["shard_comp"][:1 + 3 * W].copy(), pending=pend,
                                 want_fallback=want_fb)

    def head(self, scan, k):
        return scan.matches[:k]

    def continue_starts(self, records):
        starts, upd = [], []
        NW = self.prog["NW"]
        for nxt, R in records:
            Rv = int.from_bytes(np.asarray(R[:NW], dtype=np.uint32).tobytes(), "little")
            j, found = nxt, -1
            while j >= self.base - self.halo:
                c = self.text[j]
                x = Rv & self.prog["B"][c]
                if x & self.prog["first"]:
                    found = j
                    break
                Rv = apply_follow(self.prog, x, rev=True)
                j -= 1
                if not Rv:
                    found = -2
                    break
            starts.append(found)
            Rw = np.zeros(32, dtype=np.uint32)
            Rw[:NW] = np.frombuffer(Rv.to_bytes(4 * NW, "little"), dtype=np.uint32)
            upd.append((j, Rw))
        return starts, upd


class ThreadComm:
    """TEST INFRASTRUCTURE: allgather between threads of one process."""

    def __init__(self, world):
        self.world, self.bar, self.slots = world, threading.Barrier(world), [None] * world

    def view(self, rank):
        comm = self

        class V:
            def allgather(self, arr):
                comm.slots[rank] = np.array(arr, copy=True)
                comm.bar.wait()
                out = np.stack(comm.slots)
                comm.bar.wait()
                return out

            def allgather_obj(self, obj):
                comm.slots[rank] = obj
                comm.bar.wait()
                out = list(comm.slots)
                comm.bar.wait()
                return out
        return V()


@pytest.fixture(autouse=True, params=["thread-owned", "row-bounded"])
def mode(request):
    """Every test of this module runs with the compiler free to choose the thread-owned candidate scan, and with
    it forbidden (RRX_COMPILE_NO_DONLY)."""
    EmuBackend.DONLY = request.param == "thread-owned"
    yield request.param
    EmuBackend.DONLY = True


def run_world(rrx, pattern, text, world, halo, cuts=None):
    n = len(text)
    cuts = cuts or [((n * r // world) // 128) * 128 for r in range(world)] + [n]
    comm = ThreadComm(world)
    out, errs = [None] * world, []
    emu_api._lib()          # load the libraries before the threads race to do it
    rrx.lib()

    def work(rank):
        try:
            be = EmuBackend(rrx, pattern, text, cuts[rank], cuts[rank + 1], halo)
            res = sharded.scan_sharded(be.regex, be, comm.view(rank), rank, world, fetch=True)
            out[rank] = (res.to_host(be), res.rescans)
        except Exception as e:          # pragma: no cover
            errs.append(e)
            comm.bar.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return np.concatenate([o[0] for o in out]), [o[1] for o in out]


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("cfg_id", [1, 2, 3, 5])
def test_sharded_equals_whole(rrx, oracle_api, cfg_id, world):
    cfg = CONFIGS[cfg_id]
    text = cpu_corpus(cfg, 0, 6 * 32768 + 640)
    want = oracle_api.Oracle(cfg.pattern).lazy_matches(text)
    got, _ = run_world(rrx, cfg.pattern, text, world, halo=4096)
    assert np.array_equal(got, want)


def test_single_z_at_offset_zero_only_q_in_last_shard(rrx, oracle_api, monkeypatch):
    """the ".*" state must be carried across every shard, and the start scan of
    the one match has to travel back through all shards to offset 0."""
    n = 4 * 16384
    text = np.full(n, ord("a"), dtype=np.uint8)
    text[0] = ord("z")
    text[-5] = ord("q")
    want = oracle_api.Oracle("(x|y)*z.*q").lazy_matches(text)
    assert want.tolist() == [[0, n - 4]]
    monkeypatch.setattr(sharded, "PREFIX_WINDOW", 4096)      # shards are 16-32 KiB here: force the "whole shard" path
    for world in (2, 4):
        got, rescans = run_world(rrx, "(x|y)*z.*q", text, world, halo=256)
        assert np.array_equal(got, want), world
        assert rescans[-1] == 2          # the entry's influence never dies: last shard rescanned whole


def test_matches_straddling_a_shard_boundary(rrx, oracle_api):
    pat = "[a-z]+@[a-z]+\\.(com|org|net)"
    o = oracle_api.Oracle(pat)
    addr = b"someone@example.org"
    n = 2 * 8192
    for shift in range(len(addr) + 2):
        text = np.full(n, ord(" "), dtype=np.uint8)
        s = 8192 - shift
        text[s:s + len(addr)] = np.frombuffer(addr, dtype=np.uint8)
        got, _ = run_world(rrx, pat, text, 2, halo=128, cuts=[0, 8192, n])
        assert np.array_equal(got, o.lazy_matches(text)), shift


def test_random_texts_random_cuts(rrx, oracle_api):
    rnd = random.Random(31)
    for pat, alpha in [("(x|y)*z.*q", b"xyzq a\n"), ("a{1,100}b", b"aaaab "), ("(ab|cd)*e", b"abcde "), ("(a|b)*a(a|b){3}c", b"abc")]:
        o = oracle_api.Oracle(pat)
        for _ in range(3):
            n = rnd.randint(3000, 30000)
            text = np.frombuffer(bytes(rnd.choices(alpha, k=n)), dtype=np.uint8)
            world = rnd.randint(2, 4)
            cuts = sorted(rnd.sample(range(128, n, 128), world - 1))
            got, _ = run_world(rrx, pat, text, world, halo=rnd.choice([0, 128, 1024]), cuts=[0] + cuts + [n])
            assert np.array_equal(got, o.lazy_matches(text)), (pat, n, cuts)


def test_unaligned_cuts_with_static_closed_patterns(rrx, oracle_api):
    """Shard lengths that are not multiples of the 128-byte row: the partial last row of a shard must
    not be taken as closed by construction (acyclic automaton), or matches crossing the cut are lost."""
    pat = "abcdefgh"
    text = np.full(400, ord(" "), dtype=np.uint8)
    text[126:134] = np.frombuffer(b"abcdefgh", dtype=np.uint8)
    for cut in (127, 129, 131, 133):
        got, _ = run_world(rrx, pat, text, 2, halo=64, cuts=[0, cut, 400])
        assert got.tolist() == [[126, 134]], cut
    rnd = random.Random(77)
    for pat, alpha in [("abcab", b"abc "), ("a{1,100}b", b"aaaab "), ("(ab|cd)e(f|g)h", b"abcdefgh"), ("[a-z]+@[a-z]+\\.(com|org|net)", b"ab@.comrgnet ")]:
        o = oracle_api.Oracle(pat)
        for _ in range(2):
            n = rnd.randint(1500, 4000)
            text = np.frombuffer(bytes(rnd.choices(alpha, k=n)), dtype=np.uint8)
            world = rnd.randint(2, 4)
            cuts = sorted(rnd.sample([c for c in range(100, n - 100) if c % 128 in (1, 3, 64, 127)], world - 1))
            got, _ = run_world(rrx, pat, text, world, halo=rnd.choice([0, 128, 1024]), cuts=[0] + cuts + [n])
            assert np.array_equal(got, o.lazy_matches(text)), (pat, n, cuts)


def test_run_alive_at_the_end_of_a_full_last_row(rrx, oracle_api):
    """A shard that ends on a row boundary inside a tile, with a run alive at its last byte: the lane must not follow
    the run into the (non-existent) next row — the state at the end of the text is the shard's exit (fuzz_emu find)."""
    rnd = random.Random(5)
    for pat, alpha, wts in [("a.", b"abc\xe9\x80", [200, 1, 1, 1, 200]), ("y[a-z]+q", b"abcxyzq ", [40, 1, 5, 200, 40, 1, 200, 1]),
                            ("[a-z]+z[a-z]+", b"abcxyzq ", [1, 5, 5, 200, 1, 5, 200, 5]), ("(b[a-z]+)*", b"abc\n", [200, 200, 40, 5])]:
        o = oracle_api.Oracle(pat)
        for cut in (128, 1920):
            for _ in range(2):
                text = np.frombuffer(bytes(rnd.choices(alpha, weights=wts, k=2600)), dtype=np.uint8)
                got, _ = run_world(rrx, pat, text, 2, halo=rnd.choice([0, 128, cut]), cuts=[0, cut, 2600])
                assert np.array_equal(got, o.lazy_matches(text)), (pat, cut)


def _gloo_worker(rank, world, port, pattern, n, seed, q):
    import torch.distributed as dist
    import roaringregex_b200 as rrx
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = CONFIGS[5]
    text = cpu_corpus(cfg, 0, n)
    cuts = [((n * r // world) // 128) * 128 for r in range(world)] + [n]
    be = EmuBackend(rrx, pattern, text, cuts[rank], cuts[rank + 1], 1024)
    res = sharded.scan_sharded(be.regex, be, sharded.TorchComm(dist, "cpu"), rank, world, fetch=True)
    q.put((rank, res.to_host(be)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_over_gloo(oracle_api):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, port = 3 * 32768, 29511 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, CONFIGS[5].pattern, n, 5, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=180) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    want = oracle_api.Oracle(CONFIGS[5].pattern).lazy_matches(cpu_corpus(CONFIGS[5], 0, n))
    assert np.array_equal(np.concatenate([got[0], got[1]]), want)


def test_fold_entries_equals_fold_entry(rrx):
    """rrx_fold_entries (all ranks in one call) against rrx_fold_entry rank by rank, on random
    closed / rank-1 / identity / opaque summaries."""
    rnd = random.Random(21)
    for pat in ("(x|y)*z.*q", "a{1,100}b", "[a-z]+@[a-z]+\\.(com|org|net)"):
        r = rrx.RRegex(pat)
        w = r.info["words"]
        for _ in range(50):
            world = rnd.randint(1, 8)
            s = np.zeros((world, 1 + 3 * w), dtype=np.uint32)
            for k in range(world):
                kind = rnd.choice([0, 0, 1, 2, 0x100])          # closed, rank-1, identity, opaque closed
                s[k, 0] = kind
                s[k, 1:] = [rnd.getrandbits(32) & rnd.getrandbits(32) for _ in range(3 * w)]
                if kind & 0xff != 1:
                    s[k, 1 + w:] = 0
            both = sharded.fold_entries(r, s)
            for k in range(world):
                e, ex = sharded.fold_entry(r, s, k)
                assert np.array_equal(both[k][0], e) and both[k][1] == ex
