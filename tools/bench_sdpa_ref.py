"""The reference's attention operator on this box: F.scaled_dot_product_attention (what AttnProcessor2_0 calls,
reference utils/model_utils.py:66) at the C3 shapes, fp16 and bf16.  GPU box: python tools/bench_sdpa_ref.py"""
import torch
import torch.nn.functional as F


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
    shapes = {"self ds1 T=47520 d=40": (2, 8, 47520, 47520, 40), "self yt T=11520 d=40": (2, 8, 11520, 11520, 40),
              "self ds2 T=11880 d=80": (2, 8, 11880, 11880, 80), "cross ds1 n=14400 L=154 d=40": (8, 8, 14400, 154, 40),
              "cross yt n=5760 L=77 d=40": (8, 8, 5760, 77, 40)}
    for dt in (torch.bfloat16, torch.float16):
        for name, (B, H, T, Tk, d) in shapes.items():
            q = torch.randn(B, H, T, d, device=dev, dtype=dt)
            k = torch.randn(B, H, Tk, d, device=dev, dtype=dt)
            v = torch.randn(B, H, Tk, d, device=dev, dtype=dt)
            fl = 4.0 * B * H * T * Tk * d
            t = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
            print(f"SDPA {str(dt)[6:]:9s} {name:32s}: {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
