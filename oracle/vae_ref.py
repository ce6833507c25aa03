"""TEST INFRASTRUCTURE — not product code.

Restatement of diffusers 0.32.1 ``AutoencoderKL`` for the SD-1.5 VAE (the object behind ``pipe.vae`` at the
reference's call sites utils/VidToMe/generate_utils.py:140-172 and invert.py:118-149):
block_out_channels (128, 256, 512, 512), 2 resnet layers per block, 4 latent channels, 32 norm groups, eps 1e-6,
one single-head attention in each mid block, encoder down-sampling by stride-2 3x3 convs with (0,1,0,1) zero
padding, decoder up-sampling by nearest x2 + 3x3 conv, 1x1 quant / post_quant convs.

diffusers is neither vendored nor installable here, so this file is written from the published architecture
(state-dict key names follow diffusers so that real checkpoints load): **parity unpinned** for the third-party
arithmetic; the reference's own wrappers (scale 0.18215, 2x-1 / x/2+0.5, clamp, batching) are exercised on top.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block (residual, GroupNorm in front)."""

    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        a = torch.softmax(q @ k.transpose(1, 2) / (c ** 0.5), dim=-1) @ v
        return x + self.to_out[0](a).transpose(1, 2).reshape(b, c, h, w)


class _Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Attention(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Down(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])
        self.downsamplers = nn.ModuleList([nn.Module()]) if down else None
        if down:
            self.downsamplers[0].conv = nn.Conv2d(cout, cout, 3, stride=2, padding=0)

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))
        return x


class _Up(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(3)])
        self.upsamplers = nn.ModuleList([nn.Module()]) if up else None
        if up:
            self.upsamplers[0].conv = nn.Conv2d(cout, cout, 3, padding=1)

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class Encoder(nn.Module):
    def __init__(self, boc, in_ch=3, latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(in_ch, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(_Down(c, co, down=i < len(boc) - 1))
            c = co
        self.mid_block = _Mid(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, boc, out_ch=3, latent=4):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0])
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(_Up(c, co, up=i < len(rev) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, out_ch, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _Dist:
    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)


class AutoencoderKL(nn.Module):
    def __init__(self, block_out_channels=(128, 256, 512, 512), latent_channels=4):
        super().__init__()
        boc = tuple(block_out_channels)
        self.encoder = Encoder(boc, latent=latent_channels)
        self.decoder = Decoder(boc, latent=latent_channels)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def encode(self, x):
        return type("EncOut", (), {"latent_dist": _Dist(self.quant_conv(self.encoder(x)))})()

    def decode(self, z):
        return type("DecOut", (), {"sample": self.decoder(self.post_quant_conv(z))})()


def make_vae(seed: int = 0, dtype=torch.float32, **kw) -> AutoencoderKL:
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = AutoencoderKL(**kw)
    torch.random.set_rng_state(g)
    return m.to(dtype).eval()


# the reference's wrappers (generate_utils.py:140-172), without autocast
@torch.no_grad()
def decode_latents(vae, latents):
    imgs = vae.decode(1 / 0.18215 * latents).sample
    return (imgs / 2 + 0.5).clamp(0, 1)


@torch.no_grad()
def encode_imgs(vae, imgs):
    return vae.encode(2 * imgs - 1).latent_dist.mean * 0.18215
