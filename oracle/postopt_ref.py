"""TEST INFRASTRUCTURE — torch-autograd restatement of TC-Light's two-stage optimiser, written as
plain functions over explicit tensors and an explicit list of index batches:

  warp_bicubic        reference utils/flow_utils.py:5-16
  gauss_blur / ms_ssim_relaxed   reference utils/loss_utils.py:73-211 (+ pytorch_msssim's
                      gaussian_filter / _fspecial_gauss_1d, restated from the published algorithm:
                      PARITY UNPINNED for that third-party piece, see oracle/refshim.py)
  tv_loss             reference utils/loss_utils.py:324-339
  stage1_exposure     reference generate.py:354-451
  stage2_uvt          reference generate.py:453-533 (torch_scatter.scatter(mean) restated)

Pinned to the reference's own methods by tests/test_oracle_vs_reference.py (same seeds =>
matching loss curves and images on CPU) and by tests/golden/postopt_*.pt.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SH_C0 = 0.28209479177387814
MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def draw_batches(n_frames: int, batch_size: int, epochs: int) -> List[List[List[int]]]:
    """The batches a DataLoader(shuffle=True) yields, epoch by epoch (consumes the global CPU RNG
    like the reference's loader)."""
    loader = torch.utils.data.DataLoader(range(n_frames), batch_size=batch_size, shuffle=True)
    return [[[int(i) for i in b] for b in loader] for _ in range(epochs)]


def warp_bicubic(frames: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    n, _, h, w = frames.shape
    gx = flow[:, 0] + torch.arange(w, device=flow.device, dtype=flow.dtype)
    gy = flow[:, 1] + torch.arange(h, device=flow.device, dtype=flow.dtype)[:, None]
    gx = (gx / (w - 1) - 0.5) * 2
    gy = (gy / (h - 1) - 0.5) * 2
    return F.grid_sample(frames, torch.stack([gx, gy], dim=-1), mode="bicubic", padding_mode="zeros", align_corners=True)


def gauss_window(size=11, sigma=1.5, device=None, dtype=torch.float32):
    c = torch.arange(size, dtype=torch.float) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).to(device=device, dtype=dtype)


def gauss_blur(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    ch = x.shape[1]
    if x.shape[2] >= g.numel():
        x = F.conv2d(x, g.view(1, 1, -1, 1).repeat(ch, 1, 1, 1), groups=ch)
    if x.shape[3] >= g.numel():
        x = F.conv2d(x, g.view(1, 1, 1, -1).repeat(ch, 1, 1, 1), groups=ch)
    return x


def _ssim_pair(x, y, g, data_range=1.0, k1=0.01, k2=0.03):
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    mu1, mu2 = gauss_blur(x, g), gauss_blur(y, g)
    s11 = gauss_blur(x * x, g) - mu1.pow(2)
    s22 = gauss_blur(y * y, g) - mu2.pow(2)
    s12 = gauss_blur(x * y, g) - mu1 * mu2
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    ss = (2 * mu1 * mu2 + c1) / (mu1.pow(2) + mu2.pow(2) + c1) * cs
    return ss.flatten(2).mean(-1), cs.flatten(2).mean(-1)


def ms_ssim_relaxed(x, y, start_level=1, data_range=1.0):
    g = gauss_window(device=x.device, dtype=x.dtype)
    w = x.new_tensor(MS_WEIGHTS)
    terms = []
    for lvl in range(5):
        if lvl >= start_level:
            ss, cs = _ssim_pair(x, y, g, data_range)
        else:
            ss = cs = torch.ones_like(x[:, :, 0, 0])
        if lvl < 4:
            terms.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            # .contiguous(): torch 2.11 CUDA avg_pool2d BACKWARD is wrong for channels_last-strided
            # inputs with padding (differs from its own CPU and contiguous-CUDA results, see
            # tools/debug_pool.py); the reference's images are NHWC-strided views, so on a GPU it
            # would hit that library bug whenever a pyramid level has an odd size.  The oracle pins
            # the documented (CPU) semantics.
            x = F.avg_pool2d(x.contiguous(), kernel_size=2, padding=pad)
            y = F.avg_pool2d(y.contiguous(), kernel_size=2, padding=pad)
    terms.append(torch.relu(ss))
    stack = torch.stack(terms, dim=0)
    return torch.prod(stack ** w.view(-1, 1, 1), dim=0).mean()


def tv_loss(x, weight):
    b, c, h, w = x.shape
    ch, cw = c * (h - 1) * w, c * h * (w - 1)
    dh = (x[:, :, 1:, :] - x[:, :, :-1, :]).pow(2).sum()
    dw = (x[:, :, :, 1:] - x[:, :, :, :-1]).pow(2).sum()
    return weight * 2 * (dh / ch + dw / cw) / b


def expon_lr(step, lr_init, lr_final, max_steps):
    t = np.clip(step / max_steps, 0, 1)
    return float(np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t))


def _flow_term(images, pre_images, flows, masks, idx_t):
    warped = warp_bicubic(pre_images, flows)
    valid = idx_t > 0
    return (warped[valid] * masks[valid] - images[valid] * masks[valid]).abs().mean()


def stage1_exposure(edited, flows, masks, batches, lambda_dssim=0.2, lambda_flow=0.8, lr_init=0.01, lr_final=0.001):
    """Returns (aligned images, exposure, losses).  ``batches[epoch][i]`` = frame indices."""
    n, _, h, w = edited.shape
    bo = max(len(b) for ep in batches for b in ep)
    total = len(batches) * n // bo
    expo = torch.eye(3, 4, device=edited.device)[None].repeat(n, 1, 1).requires_grad_(True)
    opt = torch.optim.Adam([expo])
    losses = []
    for ep, epoch in enumerate(batches):
        for i, idx in enumerate(epoch):
            for gparam in opt.param_groups:
                gparam["lr"] = expon_lr(ep * n // bo + i + 1, lr_init, lr_final, total)
            idx_t = torch.tensor(idx, device=edited.device)
            both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
            pix = edited[both].permute(0, 2, 3, 1).reshape(len(both), h * w, 3)
            out = torch.bmm(pix, expo[both, :3, :3]) + expo[both, None, :3, 3]
            out = out.clamp(0, 1).reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
            img, pre = out[:len(idx)], out[len(idx):]
            tgt = edited[idx_t]
            photo = (img - tgt).abs().mean() * (1 - lambda_dssim) + (1 - ms_ssim_relaxed(img, tgt)) * lambda_dssim
            flow = _flow_term(img, pre, flows[idx_t], masks[idx_t], idx_t)
            loss = (1 - lambda_flow) * photo + lambda_flow * flow
            losses.append(loss.item())
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    with torch.no_grad():
        flat = edited.permute(0, 2, 3, 1).reshape(n, h * w, 3)
        aligned = (torch.bmm(flat, expo[:, :3, :3]) + expo[:, None, :3, 3]).clamp(0, 1)
        aligned = aligned.reshape(n, h, w, 3).permute(0, 3, 1, 2).contiguous()
    return aligned, expo.detach(), losses


def scatter_mean(values, index, size):
    out = torch.zeros((size, values.shape[1]), dtype=values.dtype, device=values.device)
    out.index_add_(0, index, values)
    cnt = torch.zeros(size, dtype=values.dtype, device=values.device)
    cnt.index_add_(0, index, torch.ones_like(index, dtype=values.dtype))
    return out / cnt.clamp_min(1)[:, None]


def stage2_uvt(edited, flows, masks, unq_inv, batches, lambda_dssim=0.2, lambda_flow=0.8, lambda_tv=0.05, feature_lr=0.05):
    """Returns (images, features_dc, losses)."""
    n, _, h, w = edited.shape
    bo = max(len(b) for ep in batches for b in ep)
    lr = feature_lr * bo / n
    unq_inv = unq_inv.long()
    size = int(unq_inv.max().item()) + 1
    with torch.no_grad():
        mean_rgb = scatter_mean(edited.permute(0, 2, 3, 1).reshape(n * h * w, 3), unq_inv, size)
        init = (mean_rgb - 0.5) / SH_C0
    fdc = init.contiguous().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [fdc], "lr": lr}], lr=0.0, eps=1e-15)
    table = unq_inv.reshape(n, h, w)
    losses = []
    for epoch in batches:
        for idx in epoch:
            idx_t = torch.tensor(idx, device=edited.device)
            both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
            ids = table[both].reshape(-1)
            rgb = torch.index_select(fdc * SH_C0 + 0.5, 0, ids).clamp(0, 1)
            out = rgb.reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
            img, pre = out[:len(idx)], out[len(idx):]
            flow = _flow_term(img, pre, flows[idx_t], masks[idx_t], idx_t)
            photo = (1 - ms_ssim_relaxed(img, edited[idx_t])) * lambda_dssim
            loss = (1 - lambda_flow) * photo + lambda_flow * flow + tv_loss(img, lambda_tv)
            losses.append(loss.item())
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    with torch.no_grad():
        images = (fdc * SH_C0 + 0.5)[unq_inv].reshape(n, h * w, 3).clamp(0, 1).reshape(n, h, w, 3).permute(0, 3, 1, 2).contiguous()
    return images, fdc.detach(), losses


def synthetic_clip(n=6, h=192, w=200, seed=0, device="cpu"):
    """Seeded synthetic stage-1/2 inputs: a smooth texture translated by a constant backward flow
    plus noise; soft masks in [0,1]; flow-tracked unique ids (integer translation => exact
    correspondences, fresh ids for pixels entering the frame)."""
    g = torch.Generator().manual_seed(seed)
    dx, dy = 2, 1                                  # content moves by (+dx, +dy) px per frame
    big = F.interpolate(torch.rand(1, 3, (h + n * dy) // 8 + 2, (w + n * dx) // 8 + 2, generator=g), size=(h + n * dy, w + n * dx),
                        mode="bicubic", align_corners=False).clamp(0, 1)[0]
    frames, ids = [], []
    big_ids = torch.arange((h + n * dy) * (w + n * dx)).reshape(h + n * dy, w + n * dx)
    for i in range(n):
        oy, ox = (n - 1 - i) * dy, (n - 1 - i) * dx
        frames.append(big[:, oy:oy + h, ox:ox + w])
        ids.append(big_ids[oy:oy + h, ox:ox + w])
    clean = torch.stack(frames)
    edited = (clean * 0.8 + 0.1 + 0.02 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    # per-frame exposure drift so stage 1 has something to fix
    gain = 1.0 + 0.05 * torch.randn(n, 1, 1, 1, generator=g)
    edited = (edited * gain).clamp(0, 1)
    past = torch.zeros(n, 2, h, w)
    past[:, 0] = -dx + 0.25 * torch.sin(torch.arange(w).float() / 17.0)[None, None, :]
    past[:, 1] = -dy + 0.25 * torch.cos(torch.arange(h).float() / 13.0)[None, :, None]
    masks = (0.5 + 0.5 * torch.rand(n, 1, h, w, generator=g))
    masks[:, :, :, :dx + 1] = 0.0
    masks[:, :, :dy + 1, :] = 0.0
    _, inv = torch.unique(torch.stack(ids).reshape(-1), return_inverse=True)
    return edited.to(device), past.to(device), masks.to(device), inv.to(device)
