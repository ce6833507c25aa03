"""TEST INFRASTRUCTURE — CPU/torch-fp32 restatement of the SD-1.5 ``UNet2DConditionModel`` as used by
TC-Light.  PARITY UNPINNED at this boundary: the arithmetic lives in diffusers==0.32.1
(reference requirements.txt:1), which is not vendored in /root/reference and not installable
offline; nothing in the reference pins it.  This file restates the published architecture
 and follows the reference's own in-repo restatements where they exist:

  * BasicTransformerBlock.forward      utils/VidToMe/vidtome/patch.py:128-201
  * Attention.forward                  utils/VidToMe/pnp_utils.py:40-97
  * ResnetBlock2D.forward              utils/VidToMe/pnp_utils.py:110-164
  * block indexing / key names         utils/VidToMe/pnp_utils.py:12-37, 100-105, 168-171
  * IC-Light 8-channel conv_in + concat_conds hook   utils/model_utils.py:21-26, 35-43

Module / parameter names equal diffusers' state-dict keys so real weights can be loaded later and
so the reference's ``vidtome.apply_patch`` (which class-swaps modules *named*
``BasicTransformerBlock`` under a ``ModelMixin``) works on it unmodified.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class ModelMixin(nn.Module):
    """Name matters: reference patch.py:263 checks ``isinstance_str(model, "ModelMixin")``."""


def timestep_embedding(timesteps: torch.Tensor, dim: int = 320, max_period: float = 10000.0) -> torch.Tensor:
    """Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0): cat([cos, sin])."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class Attention(nn.Module):
    def __init__(self, query_dim, cross_dim=None, heads=8):
        super().__init__()
        self.heads = heads
        cross_dim = query_dim if cross_dim is None else cross_dim
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(cross_dim, query_dim, bias=False)
        self.to_v = nn.Linear(cross_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        B, N, C = hidden_states.shape
        h = self.heads
        q = self.to_q(hidden_states).view(B, N, h, C // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, h, C // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, h, C // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, N, C).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.only_cross_attention = False
        self.use_ada_layer_norm = False
        self.use_ada_layer_norm_zero = False
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_dim, heads)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, timestep=None, cross_attention_kwargs=None, class_labels=None):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, cross_dim, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, cross_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, encoder_hidden_states):
        B, C, H, W = x.shape
        res = x
        h = self.proj_in(self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states=encoder_hidden_states)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(h) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_dim, heads, cross_dim, has_attn, add_down, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim, groups) for i in range(2)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim, groups) for _ in range(2)])
        self.has_attn = has_attn
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.add_down = add_down

    def forward(self, x, temb, ehs):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ehs)
            outs.append(x)
        if self.add_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb_dim, heads, cross_dim, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cross_dim, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim, groups) for _ in range(2)])

    def forward(self, x, temb, ehs):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin_list, cout, temb_dim, heads, cross_dim, has_attn, add_up, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ci, cout, temb_dim, groups) for ci in cin_list])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim, groups) for _ in cin_list])
        self.has_attn = has_attn
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x, skips, temb, ehs, upsample_size=None):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ehs)
        if self.add_up:
            x = self.upsamplers[0](x, upsample_size)
        return x


class UNet2DConditionModel(ModelMixin):
    """SD-1.5 layout: down (CrossAttn x3, Down), mid CrossAttn, up (Up, CrossAttn x3).

    ``in_channels=8`` is the IC-Light variant (latent 4 + concat_conds 4); the reference keeps
    ``config.in_channels == 4`` (model_utils.py:22-26), mirrored by ``latent_channels``.
    """

    def __init__(self, block_out_channels: Sequence[int] = (320, 640, 1280, 1280), heads: int = 8,
                 cross_attention_dim: int = 768, in_channels: int = 8, out_channels: int = 4,
                 norm_num_groups: int = 32, time_dim: Optional[int] = None):
        super().__init__()
        boc = tuple(block_out_channels)
        self.block_out_channels = boc
        self.heads = heads
        self.cross_attention_dim = cross_attention_dim
        self.in_channels = in_channels
        self.latent_channels = 4
        self.t_in = boc[0]
        temb = boc[0] * 4 if time_dim is None else time_dim
        g = norm_num_groups
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        downs, c = [], boc[0]
        for i, co in enumerate(boc):
            last = i == len(boc) - 1
            downs.append(DownBlock(c, co, temb, heads, cross_attention_dim, has_attn=not last, add_down=not last, groups=g))
            c = co
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(boc[-1], temb, heads, cross_attention_dim, g)
        # skip-channel stack in push order
        skip = [boc[0]]
        for i, co in enumerate(boc):
            skip += [co, co] + ([co] if i < len(boc) - 1 else [])
        ups, rev = [], list(reversed(boc))
        c = boc[-1]
        for i, co in enumerate(rev):
            cins = []
            for _ in range(3):
                cins.append(c + skip.pop())
                c = co
            ups.append(UpBlock(cins, co, temb, heads, cross_attention_dim, has_attn=i > 0, add_up=i < len(boc) - 1, groups=g))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.config = type("Cfg", (), {"in_channels": 4})()

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, **kw):
        # IC-Light hook (reference utils/model_utils.py:35-40)
        if cross_attention_kwargs is not None and cross_attention_kwargs.get("concat_conds") is not None:
            c_concat = cross_attention_kwargs["concat_conds"].to(sample)
            c_concat = torch.cat([c_concat] * (sample.shape[0] // c_concat.shape[0]), dim=0)
            sample = torch.cat([sample, c_concat], dim=1)
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timesteps = timestep.reshape(-1).expand(sample.shape[0]) if timestep.numel() == 1 else timestep
        t_emb = timestep_embedding(timesteps, self.t_in).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        n_up = len(self.block_out_channels) - 1
        forward_upsample_size = any(s % (2 ** n_up) != 0 for s in sample.shape[-2:])
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        for i, blk in enumerate(self.up_blocks):
            is_final = i == len(self.up_blocks) - 1
            n_res = len(blk.resnets)
            res = skips[-n_res:]
            skips = skips[:-n_res]
            up_size = skips[-1].shape[2:] if (not is_final and forward_upsample_size) else None
            x = blk(x, list(res), emb, encoder_hidden_states, up_size)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return type("UNetOut", (), {"sample": x})()


def make_unet(seed: int = 0, dtype=torch.float32, **kw) -> UNet2DConditionModel:
    """Seeded PyTorch-default init (no real weights exist offline, SURVEY.md §7 hard part 7)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = UNet2DConditionModel(**kw)
    torch.random.set_rng_state(g)
    return m.to(dtype).eval()


# ---------------------------------------------------------------------------------------------
# VidToMe-patched transformer block for the oracle UNet (reference patch.py:119-203 semantics,
# built on oracle/vidtome_ref.py so it also runs where /root/reference does not exist).
# ---------------------------------------------------------------------------------------------
class ToMeBasicTransformerBlock(BasicTransformerBlock):
    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, timestep=None, cross_attention_kwargs=None, class_labels=None):
        from . import vidtome_ref as V

        info = self._oracle_tome
        if not hasattr(self, "_state"):
            dev = hidden_states.device
            gen = torch.Generator(device=dev)
            gen.set_state(torch.cuda.get_rng_state() if dev.type == "cuda" else torch.get_rng_state())
            self._state = V.MergeState(gen)
        norm = self.norm1(hidden_states)
        merged, unmerge, _ = V.compute_merge(self._state, norm, info["size"], info["args"])
        attn = self.attn1(merged)
        hidden_states = unmerge(attn) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


def apply_oracle_patch(unet: UNet2DConditionModel, local_merge_ratio=0.6, merge_global=True, global_merge_ratio=0.5,
                       max_downsample=2, batch_size=2, align_batch=True, target_stride=4, global_rand=0.5):
    info = dict(size=None, args=dict(max_downsample=max_downsample, batch_size=batch_size, align_batch=align_batch,
                                     merge_global=merge_global, global_merge_ratio=global_merge_ratio,
                                     local_merge_ratio=local_merge_ratio, global_rand=global_rand,
                                     target_stride=target_stride))
    unet._oracle_tome = info
    unet.register_forward_pre_hook(lambda mod, a: info.__setitem__("size", (a[0].shape[2], a[0].shape[3])))
    for m in unet.modules():
        if type(m) is BasicTransformerBlock:
            m.__class__ = ToMeBasicTransformerBlock
            m._oracle_tome = info
    return unet


def reset_oracle_pool(unet: UNet2DConditionModel):
    """post_iter (reference generate_utils.py:235-238): drop every block's global-token pool."""
    for m in unet.modules():
        if hasattr(m, "_state"):
            m._state.global_tokens = None
