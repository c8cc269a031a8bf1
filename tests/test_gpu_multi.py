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
