"""Times every tcl_attention tuning variant at the C3 ds-1 shapes (bf16).  GPU box: python tools/bench_attn_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200 import ops
from tclight_b200._lib import lib


def timeit(fn, iters=3, warm=1):
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
    dt = torch.bfloat16
    shapes = {"self ds1 T=47520 d=40": (2, 8, 47520, 47520, 40, 1), "self yt T=11520 d=40": (2, 8, 11520, 11520, 40, 1),
              "self ds2 T=11880 d=80": (2, 8, 11880, 11880, 80, 1), "cross ds1 n=14400 L=154 d=40": (8, 8, 14400, 154, 40, 4), "cross yt n=5760 L=77 d=40": (8, 8, 5760, 77, 40, 4)}
    for name, (B, H, T, Tk, d, div) in shapes.items():
        dp = ops.head_pad(d)
        Tp, Tkp = (T + 7) // 8 * 8, (Tk + 7) // 8 * 8
        q = torch.randn(B, H, Tp, dp, device=dev).to(dt)
        k = torch.randn(B // div, H, Tkp, dp, device=dev).to(dt)
        vt = torch.randn(B // div, H, dp, Tkp, device=dev).to(dt)
        out = torch.empty(B, T, H * d, device=dev, dtype=dt)
        fl = 4.0 * B * H * T * Tk * d
        for var in ([0, 1, 2, 3, 4, 5, 6, 7, 8] if d == 40 else [0, 1, 3]):
            lib.tcl_debug_attention_variant(var)
            t = timeit(lambda: ops.attention(q, k, vt, T, Tk, d, kv_batch_div=div, out=out))
            print(f"{name:32s} variant {var}: {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} TFLOP/s", flush=True)
    lib.tcl_debug_attention_variant(0)


if __name__ == "__main__":
    main()
