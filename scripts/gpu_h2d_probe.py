"""Measures pinned-host -> device copy rates on the box (1-D and the strided 2-D pattern of run_streams)."""
import time, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
n = 1536 * 125 * 4096 * 2
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print("1-D H2D %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))
h2 = h.view(4096, -1); d2 = d.view(4096, -1)
w = 20 * 3072
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(6):
        d2[:, k * w:(k + 1) * w].copy_(h2[:, k * w:(k + 1) * w], non_blocking=True)
    torch.cuda.synchronize()
    print("2-D H2D (4096 rows x 61 KB, 6 windows) %.1f GB/s" % (6 * w * 4096 / (time.perf_counter() - t0) / 1e9))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print("1-D D2H %.1f GB/s" % (n / (time.perf_counter() - t0) / 1e9))
