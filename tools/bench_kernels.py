"""Kernel micro-benchmarks at the C3 (720x1280, 4-frame chunk) shapes: prints achieved TFLOP/s.
Usage (GPU box): python tools/bench_kernels.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200 import ops, _lib as L


def timeit(fn, iters=10, warm=3):
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


def main():
    dev = torch.device("cuda")
    dt = torch.float16
    print("== igemm ==")
    for name, (n, h, w, ci, co, taps) in {
        "conv3x3 ds1 320->320 (8x90x160)": (8, 90, 160, 320, 320, 9),
        "conv3x3 ds2 640->640 (8x45x80)": (8, 45, 80, 640, 640, 9),
        "conv3x3 ds4 1280->1280 (8x23x40)": (8, 23, 40, 1280, 1280, 9),
        "conv3x3 up3 960->320 (8x90x160)": (8, 90, 160, 960, 320, 9),
        "linear 115200x320 -> 960 (qkv)": (1, 1, 115200, 320, 960, 1),
        "linear 115200x320 -> 2560 (geglu)": (1, 1, 115200, 320, 2560, 1),
        "linear 115200x1280 -> 320 (ff out)": (1, 1, 115200, 1280, 320, 1),
        "linear 8192x8192x8192": (1, 1, 8192, 8192, 8192, 1),
    }.items():
        x = torch.randn(n, h, w, ci, device=dev).to(dt)
        wt = (torch.randn(co, taps * ci, device=dev) * 0.02).to(dt)
        out = torch.empty(n, h, w, co, device=dev, dtype=dt)
        t = timeit(lambda: ops.igemm([(x, taps, 1)], wt, (n, h, w), out=out))
        fl = 2.0 * n * h * w * ci * co * taps
        print(f"{name:45s} {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} TFLOP/s")
    print("== attention ==")
    for name, (B, H, T, Tk, d, div) in {
        "self ds1 merged T=47520 d=40": (2, 8, 47520, 47520, 40, 1),
        "self ds1 first-chunk T=31680 d=40": (2, 8, 31680, 31680, 40, 1),
        "self ds2 merged T=11880 d=80": (2, 8, 11880, 11880, 80, 1),
        "self ds4 T=920 d=160 (x8 img)": (8, 8, 920, 920, 160, 1),
        "cross ds1 n=14400 L=154 d=40": (8, 8, 14400, 154, 40, 4),
    }.items():
        dp = ops.head_pad(d)
        Tp = (T + 7) // 8 * 8
        Tkp = (Tk + 7) // 8 * 8
        q = torch.randn(B, H, Tp, dp, device=dev).to(dt)
        k = torch.randn(B // div, H, Tkp, dp, device=dev).to(dt)
        vt = torch.randn(B // div, H, dp, Tkp, device=dev).to(dt)
        out = torch.empty(B, T, H * d, device=dev, dtype=dt)
        t = timeit(lambda: ops.attention(q, k, vt, T, Tk, d, kv_batch_div=div, out=out), iters=3, warm=1)
        fl = 4.0 * B * H * T * Tk * d
        print(f"{name:45s} {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} TFLOP/s (algorithmic)")
        if "47520" in name or "11880" in name:
            qq = q[:, :, :T, :d].contiguous(); kk = k[:, :, :Tk, :d].contiguous(); vv = vt[:, :, :d, :Tk].transpose(2, 3).contiguous()
            t2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv), iters=3, warm=1)
            print(f"{'   torch SDPA (library) same shape':45s} {t2*1e3:8.3f} ms  {fl/t2/1e12:7.1f} TFLOP/s")
    a = torch.randn(8192, 8192, device=dev).to(dt); b = torch.randn(8192, 8192, device=dev).to(dt)
    t = timeit(lambda: a @ b.t())
    print(f"{'cuBLAS 8192^3 (library, for context)':45s} {t*1e3:8.3f} ms  {2*8192**3/t/1e12:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
