"""Sharded scans over REAL GPUs with NCCL (world size 2): the fused scan + all-gather step of
roaringregex_b200/sharded.py against the oracle.  Skipped on boxes with fewer than two GPUs (the CPU
suite covers the same logic over gloo in tests/test_sharded.py)."""
import os

import numpy as np
import pytest

from conftest import cpu_corpus
from roaringregex_b200.configs import CONFIGS

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, cfg_id, n, q):
    import torch
    import torch.distributed as dist
    import roaringregex_b200 as rrx
    from roaringregex_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = CONFIGS[cfg_id]
    regex = rrx.RRegex(cfg.pattern)
    size = n // world
    off = rank * size
    halo = min(off, 4096)
    d_buf = torch.empty(halo + size + 256, dtype=torch.uint8, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    if cfg.mode == "iid":
        rrx.corpus_iid(d_buf.data_ptr(), off - halo, halo + size, cfg.seed, cfg.alpha, sh)
    else:
        rrx.corpus_records(d_buf.data_ptr(), off - halo, halo + size, cfg.seed, sh)
    torch.cuda.synchronize()
    be = sharded.DeviceBackend(regex, d_buf[halo:].data_ptr(), size, off, halo, size // 16 + 1024, sh)
    comm = sharded.TorchComm(dist, torch.device("cuda", rank))
    res = None
    for _ in range(3):                                   # repeated steps reuse every buffer
        res = sharded.scan_sharded(regex, be, comm, rank, world, fetch=True)
    first = res.to_host(be)
    # software-pipelined steps (begin k+1 before finishing k) must give the same list every time
    prev = None
    for _ in range(4):
        tok = sharded.scan_sharded_begin(regex, be, comm)
        assert tok is not None
        if prev is not None:
            r = sharded.scan_sharded_finish(regex, be, comm, rank, world, prev, fetch=True)
            assert np.array_equal(r.to_host(be), first)
        prev = tok
    res = sharded.scan_sharded_finish(regex, be, comm, rank, world, prev, fetch=True)
    q.put((rank, res.to_host(be), res.rescans))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg_id", [2, 5, 1])
def test_two_gpus_over_nccl(oracle_api, cfg_id):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, port = 8 << 20, 29611 + os.getpid() % 200 + cfg_id
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg_id, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = {}
    for _ in range(2):
        rank, m, rescans = q.get(timeout=300)
        got[rank] = m
    [p.join(timeout=60) for p in procs]
    want = oracle_api.Oracle(CONFIGS[cfg_id].pattern).lazy_matches(cpu_corpus(CONFIGS[cfg_id], 0, n))
    assert np.array_equal(np.concatenate([got[0], got[1]]), want)


def _worker_c_abi(rank, world, port, cfg_id, n, old_mode, q):
    """The same sharded scans through the C ABI (rrx_sharded_*, csrc/rrx_dist.cu): NCCL all-gather and the fold on the
    device, repair rounds and start hand-over in C++."""
    import torch
    import torch.distributed as dist
    import roaringregex_b200 as rrx
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def bcast(raw):
        t = torch.tensor(list(raw), dtype=torch.uint8)
        dist.broadcast(t, 0)
        return bytes(t.tolist())
    comm = rrx.NcclComm(rank, world, bcast)
    cfg = CONFIGS[cfg_id]
    regex = rrx.RRegex(cfg.pattern, rrx.COMPILE_NO_DONLY if old_mode else 0)
    size = (n // world) & ~127
    off = rank * size
    halo = min(off, 256 if old_mode else 4096)
    d_buf = torch.empty(halo + size + 256, dtype=torch.uint8, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    rrx.corpus_device(cfg, d_buf.data_ptr(), off - halo, halo + size, sh)
    torch.cuda.synchronize()
    ss = rrx.ShardedScanner(regex, size, size // 16 + 1024, comm)
    ptr = d_buf[halo:].data_ptr()
    res = ss.scan(ptr, size, off, halo, sh)
    first = ss.matches(res, sh)
    # pipelined steps: begin k+1 before finish k
    ss.begin(ptr, size, off, halo, sh)
    for _ in range(3):
        ss.begin(ptr, size, off, halo, sh)
        r = ss.finish()
        assert np.array_equal(ss.matches(r, sh), first)
    r = ss.finish()
    assert np.array_equal(ss.matches(r, sh), first)
    q.put((rank, first, r["rescans"], ss.digest(r, stream=sh)))
    dist.barrier()
    del ss
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg_id,old_mode", [(2, False), (5, False), (5, True), (1, False), (3, True)])
def test_two_gpus_c_abi_sharded(oracle_api, cfg_id, old_mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, port = 8 << 20, 29811 + os.getpid() % 200 + cfg_id + (7 if old_mode else 0)
    procs = [ctx.Process(target=_worker_c_abi, args=(r, 2, port, cfg_id, n, old_mode, q)) for r in range(2)]
    [p.start() for p in procs]
    got, dg = {}, {}
    for _ in range(2):
        rank, m, rescans, d = q.get(timeout=300)
        got[rank], dg[rank] = m, d
    [p.join(timeout=60) for p in procs]
    text = cpu_corpus(CONFIGS[cfg_id], 0, 2 * ((n // 2) & ~127))
    want = oracle_api.Oracle(CONFIGS[cfg_id].pattern).lazy_matches(text)
    both = np.concatenate([got[0], got[1]])
    assert np.array_equal(both, want)
    # shard digests add / xor up to the digest of the whole list
    w = oracle_api.digest(want)
    assert dg[0][0] + dg[1][0] == w[0] and (dg[0][1] + dg[1][1]) % (1 << 64) == w[1] and dg[0][2] ^ dg[1][2] == w[2]
