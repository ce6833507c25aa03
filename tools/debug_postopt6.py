import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import postopt_ref as O
from tclight_b200._lib import lib, check, stream_ptr
cuda = torch.device("cuda")
g = O.gauss_window(device=cuda)
for (h, w) in [(22, 24), (22, 25), (23, 24), (22, 26), (45, 51), (44, 50), (33, 47), (90, 160), (11, 13), (12, 12)]:
    for use_ssim in (0, 1):
        torch.manual_seed(1)
        X = torch.rand(6, h, w, device=cuda).requires_grad_(True)
        Y = (X.detach() + 0.1 * torch.randn(6, h, w, device=cuda)).clamp(0, 1).contiguous()
        coef = torch.randn(6, device=cuda)
        ss, cs = O._ssim_pair(X.view(2, 3, h, w), Y.view(2, 3, h, w), g, 1.0)
        n_valid = (h - 10) * (w - 10)
        tgt = (ss if use_ssim else cs).reshape(6) * n_valid          # sum over the valid map
        (tgt * coef).sum().backward()
        sums = torch.zeros(6, 2, device=cuda); dX = torch.zeros(6, h, w, device=cuda)
        check(lib.tcl_debug_ssim_level(X.detach().contiguous().data_ptr(), Y.data_ptr(), 6, h, w, coef.data_ptr(), use_ssim, sums.data_ptr(), dX.data_ptr(), stream_ptr()), "dbg")
        rel = ((dX - X.grad).norm() / X.grad.norm()).item()
        fw = ((sums[:, 1 if use_ssim else 0] - tgt.detach()).abs().max() / tgt.detach().abs().max()).item()
        colerr = ((dX - X.grad) ** 2).sum(dim=(0, 1)).sqrt() / (X.grad ** 2).sum(dim=(0, 1)).sqrt()
        print(f"{h}x{w} use_ssim={use_ssim}: fwd rel {fw:.1e} bwd rel {rel:.2e}", "" if rel < 1e-4 else [f"{v:.0e}" for v in colerr.tolist()][:28])
