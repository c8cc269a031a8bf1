"""Where a sharded step's host time goes (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, numpy as np
import roaringregex_b200 as rrx
from roaringregex_b200 import sharded
from roaringregex_b200.configs import CONFIGS
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = CONFIGS[2]; size = 1 << 30
regex = rrx.RRegex(cfg.pattern)
sh = torch.cuda.current_stream().cuda_stream
d = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
rrx.corpus_records(d.data_ptr(), rank * size, size, cfg.seed, sh)
be = sharded.DeviceBackend(regex, d.data_ptr(), size, rank * size, 0, size // 48, sh)
comm = sharded.TorchComm(dist, torch.device("cuda", local))
for _ in range(5): sharded.scan_sharded(regex, be, comm, rank, world)
torch.cuda.synchronize(); dist.barrier()
import cProfile, pstats
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(50): sharded.scan_sharded(regex, be, comm, rank, world)
pr.disable()
dt = (time.perf_counter() - t0) / 50
if rank == 0:
    print("ms per step", dt * 1e3)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
dist.destroy_process_group()
