"""Per-term gradient check of the stage-2 CUDA iteration vs torch autograd (debug helper)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import postopt_ref as O
from tclight_b200 import postopt as P
from tclight_b200._lib import lib, check, stream_ptr

cuda = torch.device("cuda")
h, w, n = 176, 192, 5
edited, flows, masks, inv = O.synthetic_clip(n=n, h=h, w=w, seed=3, device=cuda)
ds = P.OptDataset(edited, flows, masks, device=cuda)
idx = [3, 0, 4, 1]
size = int(inv.max().item()) + 1
mean_rgb = O.scatter_mean(edited.permute(0, 2, 3, 1).reshape(-1, 3), inv, size)
torch.manual_seed(0)
for clampy in (False, True):
    fdc0 = ((mean_rgb - 0.5) / O.SH_C0 + (0.3 if clampy else 0.0) * torch.randn(size, 3, device=cuda)).contiguous()
    for name, (ld, lf, ltv) in {"flow": (0.0, 1.0, 0.0), "tv": (0.0, 0.0, 0.05), "ssim": (1.0, 0.0, 0.0), "all": (0.2, 0.8, 0.05)}.items():
        fdc = fdc0.clone().requires_grad_(True)
        idx_t = torch.tensor(idx, device=cuda)
        both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
        rgb = torch.index_select(fdc * O.SH_C0 + 0.5, 0, inv.reshape(n, h, w)[both].reshape(-1)).clamp(0, 1)
        out = rgb.reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
        img, pre = out[:4], out[4:]
        flow = O._flow_term(img, pre, flows[idx_t], masks[idx_t], idx_t)
        photo = (1 - O.ms_ssim_relaxed(img, edited[idx_t])) * ld
        loss = (1 - lf) * photo + lf * flow + O.tv_loss(img, ltv)
        loss.backward()
        ctx = P._Context(ds, ld, lf, ltv, 4)
        ids = inv.to(torch.int32).contiguous()
        p = fdc0.clone()
        g, m, v = (torch.zeros_like(p) for _ in range(3))
        lo = torch.zeros(3, device=cuda)
        arr = (C.c_int * 4)(*idx)
        check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, 4, ids.data_ptr(), size, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                    0.0, 0.9, 0.999, 1e-15, 1, lo.data_ptr(), stream_ptr()), "uvt")
        grad = m / 0.1
        rel = ((grad - fdc.grad).norm() / fdc.grad.norm().clamp_min(1e-30)).item()
        d = (grad - fdc.grad).abs()
        print(f"clamp={clampy} {name:5s} rel-L2 {rel:.3e} max|d| {d.max().item():.3e} max|g| {fdc.grad.abs().max().item():.3e} "
              f"loss {lo[0].item():.8f} vs {loss.item():.8f}  flow {lo[1].item():.8f} vs {flow.item():.8f} photo {lo[2].item():.8f} vs {photo.item():.8f}")
