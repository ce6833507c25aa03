"""HBM-bound norm kernels at the C3 ds-1 shapes: achieved GB/s (read + write bytes / time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200 import ops


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


dev = torch.device("cuda"); dt = torch.bfloat16
big = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)     # L2 flush between iterations


def flushed(fn):
    def g():
        big.fill_(1)
        fn()
    return g


t_flush = timeit(lambda: big.fill_(1))
for rows, C in [(115200, 320), (28800, 640), (7360, 1280)]:
    x = torch.randn(rows, C, device=dev).to(dt); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    out = torch.empty_like(x)
    t = timeit(flushed(lambda: ops.layernorm(x, g, b, out=out))) - t_flush
    print(f"layernorm [{rows},{C}]: {t*1e6:7.1f} us  {2*x.numel()*2/t/1e9:7.0f} GB/s")
for n, h, w, c1, c2 in [(8, 90, 160, 320, 0), (8, 90, 160, 320, 320), (8, 45, 80, 640, 640), (8, 90, 160, 640, 320)]:
    x1 = torch.randn(n, h, w, c1, device=dev).to(dt)
    x2 = torch.randn(n, h, w, c2, device=dev).to(dt) if c2 else None
    C = c1 + c2
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    t = timeit(flushed(lambda: ops.groupnorm(x1, g, b, 32, 1e-5, True, x2=x2))) - t_flush
    nbytes = n * h * w * C * 2
    print(f"groupnorm+silu [{n},{h},{w},{c1}+{c2}] (stats + apply): {t*1e6:7.1f} us  {3*nbytes/t/1e9:7.0f} GB/s (2 reads + 1 write)")
