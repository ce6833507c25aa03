import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import postopt_ref as O
from tclight_b200 import postopt as P
from tclight_b200._lib import lib, check, stream_ptr
cuda = torch.device("cuda")
n = 3
for (h, w) in [(176, 192), (176, 203), (177, 192), (178, 194), (192, 208), (176, 256), (256, 192), (200, 200), (184,192), (176,200)]:
    edited, flows, masks, _ = O.synthetic_clip(n=n, h=h, w=w, seed=3, device=cuda)
    inv = torch.arange(n * h * w, device=cuda)
    ds = P.OptDataset(edited, flows, masks, device=cuda)
    size = n * h * w
    torch.manual_seed(0)
    fdc0 = (((edited.permute(0, 2, 3, 1).reshape(-1, 3) - 0.5) / O.SH_C0) + 0.3 * torch.randn(size, 3, device=cuda)).contiguous()
    idx = [2, 1]; nb = 2
    fdc = fdc0.clone().requires_grad_(True)
    idx_t = torch.tensor(idx, device=cuda)
    both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
    rgb = torch.index_select(fdc * O.SH_C0 + 0.5, 0, inv.reshape(n, h, w)[both].reshape(-1)).clamp(0, 1)
    out = rgb.reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
    img = out[:nb]
    loss = (1 - O.ms_ssim_relaxed(img, edited[idx_t]))
    loss.backward()
    ctx = P._Context(ds, 1.0, 0.0, 0.0, nb)
    ids = inv.to(torch.int32).contiguous()
    p = fdc0.clone(); g, m, v = (torch.zeros_like(p) for _ in range(3)); lo = torch.zeros(3, device=cuda)
    arr = (C.c_int * nb)(*idx)
    check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, nb, ids.data_ptr(), size, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 0.0, 0.9, 0.999, 1e-15, 1, lo.data_ptr(), stream_ptr()), "uvt")
    grad = (m / 0.1).reshape(n, h, w, 3); ref = fdc.grad.reshape(n, h, w, 3)
    rel = ((grad - ref).norm() / ref.norm()).item()
    # where is it wrong?  error energy by 32-pixel column/row bands at level 0 (64 px = one level-1 tile)
    d2 = ((grad - ref) ** 2).sum(dim=(0, 3))
    r2 = (ref ** 2).sum(dim=(0, 3))
    colband = [f"{(d2[:, x0:x0+64].sum() / r2[:, x0:x0+64].sum().clamp_min(1e-30)).sqrt().item():.1e}" for x0 in range(0, w, 64)]
    rowband = [f"{(d2[y0:y0+64].sum() / r2[y0:y0+64].sum().clamp_min(1e-30)).sqrt().item():.1e}" for y0 in range(0, h, 64)]
    print(f"{h}x{w}: rel {rel:.2e} loss {lo[0].item():.7f}/{loss.item():.7f} col-bands {colband} row-bands {rowband}")
