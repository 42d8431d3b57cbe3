"""PCIe probe: pinned H2D / D2H copy time vs size (CUDA events), and concurrent H2D + D2H."""
import torch

dev = "cuda:0"
for nbytes in (1 << 12, 1 << 16, 393216, 1572864, 6422528, 1 << 25):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 20
        print(f"{name} {nbytes:9d} B: {us:8.2f} us  {nbytes / us / 1e3:7.2f} GB/s", flush=True)
import time
h = torch.empty(6422528, dtype=torch.uint8).pin_memory()
d = torch.empty(6422528, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
print("D2H 6.4MB + sync wall:", (time.perf_counter() - t0) / 50 * 1e6, "us")
t0 = time.perf_counter()
for _ in range(200):
    torch.cuda.synchronize()
print("bare synchronize:", (time.perf_counter() - t0) / 200 * 1e6, "us")
