"""TEST INFRASTRUCTURE — not product code.

CPU restatement (torch fp32 / integer tensor ops, written independently of the reference's loop
structure) of the producer side of stage 2:

  soft_mask_bwds   utils/flow_utils.py:40-54   get_soft_mask_bwds
  flow_ids         utils/flow_utils.py:56-92   get_flowid
  unique_inverse   utils/general_utils.py:223-256 voxelization(voxel_size=None)  (= torch.unique(dim=0,
                   return_inverse=True) on one id column)
  warp             utils/flow_utils.py:5-16    warp_flow

Pinned against the unmodified reference functions in tests/test_oracle_vs_reference.py and by the
golden file tests/golden/flowid_producer.pt (oracle/make_goldens.py).

Collision rule of get_flowid: when several pixels of frame i-1 land on the same pixel of frame i, the
reference's advanced-index assignment (`flow_ids[i, y, x] = ...`, flow_utils.py:84) keeps the LAST
writer in row-major source order on CPU (on CUDA the winner is unspecified).  The restatement — and
the CUDA kernel — define the winner as the source pixel with the largest linear index.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def warp(frames: torch.Tensor, past_flows: torch.Tensor) -> torch.Tensor:
    """Bicubic backward warp: sample `frames` at pixel + flow (zeros padding, align_corners=True)."""
    N, _, H, W = frames.shape
    xs = torch.arange(W, dtype=past_flows.dtype)[None, None, :] + past_flows[:, 0]
    ys = torch.arange(H, dtype=past_flows.dtype)[None, :, None] + past_flows[:, 1]
    gx = (xs / (W - 1) - 0.5) * 2
    gy = (ys / (H - 1) - 0.5) * 2
    grid = torch.stack([gx, gy], dim=-1)
    return F.grid_sample(frames, grid, mode="bicubic", padding_mode="zeros", align_corners=True)


def _norm2(v: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1])


def soft_mask_bwds(org_images, flows, past_flows, alpha=0.1, beta=1e2, diff_threshold=0.1):
    """[N,1,H,W] soft backward-consistency mask; frame 0 is all ones."""
    N = org_images.shape[0]
    out = torch.ones_like(org_images[:, 0])
    if N < 2:
        return out[:, None]
    thr = org_images.max().item() * diff_threshold
    pf = past_flows[1:]
    f2b = warp(flows[:-1], pf)
    err = _norm2(pf + f2b) - ((_norm2(pf) + _norm2(f2b)) + 1) * alpha
    m1 = torch.sigmoid(-beta * err)
    diff = (warp(org_images[:-1], pf) - org_images[1:]).abs().amax(dim=1)
    m2 = torch.sigmoid(-beta * (diff - thr))
    out[1:] = (out[1:] * m1) * m2
    return out[:, None]


def flow_ids(frames, flows, mask_bwds, rgb_threshold=0.01):
    """[N,H,W] int32 track ids: forward-propagated along rounded forward flow, fresh ids elsewhere."""
    N, _, H, W = frames.shape
    P = H * W
    frames = frames.to(flows.dtype)
    thr = frames.max().item() * rgb_threshold
    ids = torch.empty((N, P), dtype=torch.int64)
    ids[0] = torch.arange(P)
    last = P
    lin = torch.arange(P)
    gx = (lin % W).to(flows.dtype)
    gy = (lin // W).to(flows.dtype)
    for i in range(1, N):
        x = (gx + flows[i - 1, 0].reshape(-1)).round().to(torch.int64)
        y = (gy + flows[i - 1, 1].reshape(-1)).round().to(torch.int64)
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H) & (mask_bwds[i, 0].reshape(-1) > 0.5)
        tgt = (y * W + x).clamp(0, P - 1)
        cur = frames[i].reshape(3, P)
        prev = frames[i - 1].reshape(3, P)
        ok &= (cur[:, tgt] - prev).abs().amax(dim=0) < thr
        # winner per target = the largest source index among the valid sources
        win = torch.full((P,), -1, dtype=torch.int64)
        win.scatter_reduce_(0, tgt[ok], lin[ok], reduce="amax", include_self=True)
        has = win >= 0
        row = torch.empty(P, dtype=torch.int64)
        row[has] = ids[i - 1][win[has]]
        n_new = int((~has).sum())
        row[~has] = last + torch.arange(n_new)
        last += n_new
        ids[i] = row
    return ids.reshape(N, H, W).to(torch.int32)


def unique_inverse(ids: torch.Tensor) -> torch.Tensor:
    """Inverse map of torch.unique(ids[:, None], dim=0, return_inverse=True): rank of every id among the
    sorted distinct ids (int64)."""
    flat = ids.reshape(-1).to(torch.int64)
    present = torch.zeros(int(flat.max()) + 1, dtype=torch.int64)
    present[flat] = 1
    rank = torch.cumsum(present, 0) - 1
    return rank[flat]


def synthetic_scene(n=6, h=40, w=56, seed=0, occlude=True):
    """Seeded small clip with sub-pixel flows, an occluding square and brightness noise, so that soft masks,
    flow cuts, collisions and out-of-frame targets are all exercised.  Returns (frames [n,3,h,w] in [0,1],
    fwd flows, bwd flows)."""
    g = torch.Generator().manual_seed(seed)
    big = F.interpolate(torch.rand(1, 3, h // 4 + 8, w // 4 + 8, generator=g), size=(h + 4 * n, w + 4 * n), mode="bicubic",
                        align_corners=False)[0].clamp(0, 1)
    frames = torch.empty(n, 3, h, w)
    fwd = torch.empty(n, 2, h, w)
    bwd = torch.empty(n, 2, h, w)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    for f in range(n):
        oy, ox = 2 * (n - 1 - f), 3 * (n - 1 - f)
        frames[f] = big[:, oy:oy + h, ox:ox + w]
        if occlude:
            cy, cx = 8 + 3 * f, 10 + 2 * f
            frames[f, :, cy:cy + 9, cx:cx + 9] = torch.tensor([0.9, 0.2, 0.1])[:, None, None]
        frames[f] += 0.004 * torch.randn(3, h, w, generator=g)
        fwd[f, 0] = 3.0 + 0.6 * torch.sin(yy / 5.0) + 0.3 * torch.randn(h, w, generator=g)
        fwd[f, 1] = 2.0 + 0.6 * torch.cos(xx / 7.0) + 0.3 * torch.randn(h, w, generator=g)
        bwd[f, 0] = -3.0 - 0.6 * torch.sin(yy / 5.0) + 0.05 * torch.randn(h, w, generator=g)
        bwd[f, 1] = -2.0 - 0.6 * torch.cos(xx / 7.0) + 0.05 * torch.randn(h, w, generator=g)
    return frames.clamp(0, 1), fwd, bwd
