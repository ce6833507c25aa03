"""Launch one representative instance of each hot kernel (for `ncu --set full -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
dev = torch.device("cuda"); dt = torch.bfloat16
torch.manual_seed(0)
if which.startswith("attn"):
    var = int(which[4:] or -1)
    if var != -1:                                    # tuning variants live in libtclight_tuning.so (include/tclight_tuning.h)
        from tclight_b200 import _lib
        lib = _lib.load_tuning_lib()
        ops.lib = lib
        lib.tcl_debug_attention_variant(var)
    B, H, T, d = 2, 8, 47520, 40
    dp = ops.head_pad(d); Tp = (T + 7) // 8 * 8
    q = torch.randn(B, H, Tp, dp, device=dev).to(dt); k = torch.randn(B, H, Tp, dp, device=dev).to(dt)
    vt = torch.randn(B, H, dp, Tp, device=dev).to(dt)
    for _ in range(2):
        ops.attention(q, k, vt, T, T, d)
elif which == "conv":
    x = torch.randn(8, 90, 160, 320, device=dev).to(dt); w = (torch.randn(320, 9 * 320, device=dev) * 0.02).to(dt)
    for _ in range(3):
        ops.igemm([(x, 9, 1)], w, (8, 90, 160))
elif which == "match":
    a = torch.randn(2, 43200, 320, device=dev).to(dt); b = torch.randn(2, 14400, 320, device=dev).to(dt)
    for _ in range(2):
        ops.vidtome_match(a, b, True)
elif which == "stage2":
    from tclight_b200 import postopt
    print(postopt.bench_stage2(dev, 64, 720, 1280, iters=3))
torch.cuda.synchronize()
