import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import roaringregex_b200 as rrx
from roaringregex_b200.configs import CONFIGS
cfg = CONFIGS[2]
size = 1 << 30
regex = rrx.RRegex(cfg.pattern)
d = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
rrx.corpus_records(d.data_ptr(), 0, size, cfg.seed, 0)
torch.cuda.synchronize()
if "--scanner-first" in sys.argv:
    sc = rrx.Scanner(regex, size, size // 48)
    sc.scan(d.data_ptr(), size, 0, 0, None, 0); print("scanner result", sc.result(0)["n_matches"], "jit", sc.is_jit)
h = torch.empty(size, dtype=torch.uint8).pin_memory()
h.copy_(d[:size]); torch.cuda.synchronize()
hn = h.numpy()
t0 = time.perf_counter(); regex.scan_host(hn[:1 << 20]); print("warm 1MiB", time.perf_counter() - t0)
for i in range(5):
    t0 = time.perf_counter(); m = regex.scan_host(hn, 0); n = len(m); dt = time.perf_counter() - t0
    print(i, n, round(dt * 1e3, 2), "ms", round(size / dt / 1e9, 1), "GB/s")
    del m
