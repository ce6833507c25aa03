"""Weight repacking from diffusers state-dict layout to the kernels' layouts (host side, once)."""
from __future__ import annotations

import torch


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """[Co, Ci, 3, 3] -> [Co, 9*Ci], k = (ky*3 + kx)*Ci + ci (tap-major, matches tcl_igemm)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def interleave_geglu(w: torch.Tensor, b: torch.Tensor | None):
    """GEGLU projection [8C, K] (rows 0..4C value, 4C..8C gate) -> per 256-row tile
    [128 value rows | 128 gate rows] of the same output channels (TCL_EPI_GEGLU layout)."""
    n2 = w.shape[0]
    half = n2 // 2
    assert half % 128 == 0, "GEGLU inner width must be a multiple of 128"
    idx = []
    for t in range(half // 128):
        idx.append(torch.arange(t * 128, (t + 1) * 128))
        idx.append(half + torch.arange(t * 128, (t + 1) * 128))
    idx = torch.cat(idx).to(w.device)
    wi = w.index_select(0, idx).contiguous()
    bi = None if b is None else b.index_select(0, idx).contiguous()
    return wi, bi
