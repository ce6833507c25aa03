import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import postopt_ref as O
from tclight_b200 import postopt as P
from tclight_b200._lib import lib, check, stream_ptr
cuda = torch.device("cuda")
h, w, n = 176, 192, 5
edited, flows, masks, inv = O.synthetic_clip(n=n, h=h, w=w, seed=3, device=cuda)
inv = torch.arange(n * h * w, device=cuda)            # one row per pixel: gradient image == fdc.grad
ds = P.OptDataset(edited, flows, masks, device=cuda)
idx = [3, 1]
size = n * h * w
fdc0 = ((edited.permute(0, 2, 3, 1).reshape(-1, 3) - 0.5) / O.SH_C0).contiguous()
ld, lf, ltv = 0.0, 1.0, 0.0
fdc = fdc0.clone().requires_grad_(True)
idx_t = torch.tensor(idx, device=cuda)
both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
rgb = torch.index_select(fdc * O.SH_C0 + 0.5, 0, inv.reshape(n, h, w)[both].reshape(-1)).clamp(0, 1)
out = rgb.reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
img, pre = out[:2], out[2:]
flow = O._flow_term(img, pre, flows[idx_t], masks[idx_t], idx_t)
flow.backward()
ctx = P._Context(ds, ld, lf, ltv, 2)
ids = inv.to(torch.int32).contiguous()
p = fdc0.clone(); g, m, v = (torch.zeros_like(p) for _ in range(3)); lo = torch.zeros(3, device=cuda)
arr = (C.c_int * 2)(*idx)
check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, 2, ids.data_ptr(), size, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 0.0, 0.9, 0.999, 1e-15, 1, lo.data_ptr(), stream_ptr()), "uvt")
grad = (m / 0.1).reshape(n, h, w, 3)
ref = fdc.grad.reshape(n, h, w, 3)
for f in range(n):
    d = (grad[f] - ref[f]).abs()
    print(f"frame {f}: |ref| max {ref[f].abs().max().item():.3e}  max|d| {d.max().item():.3e} n(d>1e-9) {(d>1e-9).sum().item()} relL2 {((grad[f]-ref[f]).norm()/ref[f].norm().clamp_min(1e-30)).item():.3e}")
    if d.max() > 1e-9:
        yy, xx, cc = torch.where(d > 0.5 * d.max())
        print("   worst at (y,x,c):", list(zip(yy[:6].tolist(), xx[:6].tolist(), cc[:6].tolist())))
        bad = d > 1e-9
        ys, xs, _ = torch.where(bad)
        print("   bad y range", ys.min().item(), ys.max().item(), " x range", xs.min().item(), xs.max().item())
        k = (yy[0].item(), xx[0].item(), cc[0].item())
        print("   values mine/ref:", grad[f][k].item(), ref[f][k].item())
