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


def random_state_dict(seed: int = 0, block_out_channels=(320, 640, 1280, 1280), cross_attention_dim: int = 768,
                      in_channels: int = 8, out_channels: int = 4, scale: float = 1.0):
    """Seeded random weights with diffusers' SD-1.5 UNet state-dict keys and shapes (no real
    weights exist offline; real `realistic-vision-v51` + IC-Light offsets load through the same
    keys).  Init: uniform(+-1/sqrt(fan_in)) for conv/linear, ones/zeros for norms."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, o, i, bias=True):
        b = 1.0 / (i ** 0.5)
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b * scale
        if bias:
            sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def conv(name, o, i, k):
        b = 1.0 / ((i * k * k) ** 0.5)
        sd[name + ".weight"] = (torch.rand(o, i, k, k, generator=g) * 2 - 1) * b * scale
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def norm(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)

    boc = tuple(block_out_channels)
    temb = boc[0] * 4
    conv("conv_in", boc[0], in_channels, 3)
    lin("time_embedding.linear_1", temb, boc[0])
    lin("time_embedding.linear_2", temb, temb)

    def resnet(p, ci, co):
        norm(p + "norm1", ci)
        conv(p + "conv1", co, ci, 3)
        lin(p + "time_emb_proj", co, temb)
        norm(p + "norm2", co)
        conv(p + "conv2", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut", co, ci, 1)

    def attn(p, c):
        norm(p + "norm", c)
        conv(p + "proj_in", c, c, 1)
        t = p + "transformer_blocks.0."
        for i in (1, 2, 3):
            norm(t + f"norm{i}", c)
        for a, kd in (("attn1", c), ("attn2", cross_attention_dim)):
            lin(t + a + ".to_q", c, c, bias=False)
            lin(t + a + ".to_k", c, kd, bias=False)
            lin(t + a + ".to_v", c, kd, bias=False)
            lin(t + a + ".to_out.0", c, c)
        lin(t + "ff.net.0.proj", 8 * c, c)
        lin(t + "ff.net.2", c, 4 * c)
        conv(p + "proj_out", c, c, 1)

    skip = [boc[0]]
    c = boc[0]
    for i, co in enumerate(boc):
        last = i == len(boc) - 1
        for j in range(2):
            resnet(f"down_blocks.{i}.resnets.{j}.", c if j == 0 else co, co)
            if not last:
                attn(f"down_blocks.{i}.attentions.{j}.", co)
            skip.append(co)
        if not last:
            conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
            skip.append(co)
        c = co
    resnet("mid_block.resnets.0.", boc[-1], boc[-1])
    attn("mid_block.attentions.0.", boc[-1])
    resnet("mid_block.resnets.1.", boc[-1], boc[-1])
    c = boc[-1]
    for i, co in enumerate(reversed(boc)):
        for j in range(3):
            resnet(f"up_blocks.{i}.resnets.{j}.", c + skip.pop(), co)
            if i > 0:
                attn(f"up_blocks.{i}.attentions.{j}.", co)
            c = co
        if i < len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
    norm("conv_norm_out", boc[0])
    conv("conv_out", out_channels, boc[0], 3)
    return sd


def random_vae_state_dict(seed: int = 0, block_out_channels=(128, 256, 512, 512), latent_channels: int = 4,
                          in_channels: int = 3, out_channels: int = 3):
    """Seeded random weights with diffusers' SD-1.5 ``AutoencoderKL`` state-dict keys and shapes (what
    ``pipe.vae.state_dict()`` returns; a real checkpoint loads through the same keys)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, o, i):
        b = 1.0 / (i ** 0.5)
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def conv(name, o, i, k):
        b = 1.0 / ((i * k * k) ** 0.5)
        sd[name + ".weight"] = (torch.rand(o, i, k, k, generator=g) * 2 - 1) * b
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    def norm(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)

    def resnet(p, ci, co):
        norm(p + "norm1", ci)
        conv(p + "conv1", co, ci, 3)
        norm(p + "norm2", co)
        conv(p + "conv2", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut", co, ci, 1)

    def mid(p, c):
        resnet(p + "resnets.0.", c, c)
        a = p + "attentions.0."
        norm(a + "group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(a + n, c, c)
        resnet(p + "resnets.1.", c, c)

    boc = tuple(block_out_channels)
    conv("encoder.conv_in", boc[0], in_channels, 3)
    c = boc[0]
    for i, co in enumerate(boc):
        resnet(f"encoder.down_blocks.{i}.resnets.0.", c, co)
        resnet(f"encoder.down_blocks.{i}.resnets.1.", co, co)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        c = co
    mid("encoder.mid_block.", c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * latent_channels, c, 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    conv("post_quant_conv", latent_channels, latent_channels, 1)
    rev = list(reversed(boc))
    conv("decoder.conv_in", rev[0], latent_channels, 3)
    mid("decoder.mid_block.", rev[0])
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(3):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", c if j == 0 else co, co)
        if i < len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        c = co
    norm("decoder.conv_norm_out", c)
    conv("decoder.conv_out", out_channels, c, 3)
    return sd
