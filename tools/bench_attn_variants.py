"""Times the tcl_attention tuning variants at the C3 shapes (see the dispatch in csrc/attn.cu) and checks each against
torch SDPA on the same operands.  GPU box: python tools/bench_attn_variants.py [bf16|fp16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tclight_b200 import ops
from tclight_b200 import _lib
lib = _lib.load_tuning_lib()
assert lib is not None, "build the tuning library first: make -C tclight_b200/csrc tuning"
ops.lib = lib            # ops.attention now calls the tuning build (include/tclight_tuning.h)


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
    dt = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
    shapes = {"self ds1 T=47520 d=40": (2, 8, 47520, 47520, 40, 1), "self yt T=11520 d=40": (2, 8, 11520, 11520, 40, 1),
              "self ds2 T=11880 d=80": (2, 8, 11880, 11880, 80, 1), "self ds4 T=3680 d=160": (2, 8, 3680, 3680, 160, 1),
              "cross ds1 n=14400 L=154 d=40": (8, 8, 14400, 154, 40, 4), "cross yt n=5760 L=77 d=40": (8, 8, 5760, 77, 40, 4)}
    for name, (B, H, T, Tk, d, div) in shapes.items():
        dp = ops.head_pad(d)
        Tp, Tkp = (T + 7) // 8 * 8, (Tk + 7) // 8 * 8
        q = torch.zeros(B, H, Tp, dp, device=dev, dtype=dt)
        k = torch.zeros(B // div, H, Tkp, dp, device=dev, dtype=dt)
        vt = torch.zeros(B // div, H, dp, Tkp, device=dev, dtype=dt)
        q[..., :T, :d] = torch.randn(B, H, T, d, device=dev) * 1.5
        k[..., :Tk, :d] = torch.randn(B // div, H, Tk, d, device=dev) * 1.5
        vt[..., :d, :Tk] = torch.randn(B // div, H, d, Tk, device=dev)
        out = torch.empty(B, T, H * d, device=dev, dtype=dt)
        fl = 4.0 * B * H * T * Tk * d
        # reference on a slice of the queries (fp32 softmax through SDPA on fp32 copies)
        ns = min(T, 2048)
        kk = k[..., :Tk, :d].float().repeat_interleave(div, 0)
        vv = vt[..., :d, :Tk].float().transpose(-1, -2).repeat_interleave(div, 0)
        ref = F.scaled_dot_product_attention(q[..., :ns, :d].float(), kk, vv).permute(0, 2, 1, 3).reshape(B, ns, H * d)
        variants = ([-1, 0, 5, 8] if dt == torch.bfloat16 else [-1, 0, 8]) if (d == 40 and Tk > 256) else ([-1, 0] if d == 80 else [-1])
        for trim in (1, 0):
            lib.tcl_debug_attention_trim(trim)
            for var in variants:
                if trim == 0 and var not in (-1, 0):
                    continue
                lib.tcl_debug_attention_variant(var)
                t = timeit(lambda: ops.attention(q, k, vt, T, Tk, d, kv_batch_div=div, out=out))
                err = ((out[:, :ns].float() - ref).norm() / ref.norm()).item()
                print(f"{name:30s} trim {trim} variant {var:2d}: {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} TFLOP/s   rel-L2 vs fp32 {err:.2e}", flush=True)
    lib.tcl_debug_attention_variant(-1)
    lib.tcl_debug_attention_trim(1)


if __name__ == "__main__":
    main()
