"""SD-1.5 VAE (diffusers ``AutoencoderKL``) on the B200 kernels: the object behind ``pipe.vae`` at the reference's
call sites (utils/VidToMe/generate_utils.py:140-172 ``decode_latents / encode_imgs`` and invert.py:118-149) —
SURVEY.md §8f rank 1.

    vae = AutoencoderKLB200(pipe.vae.state_dict(), device="cuda", dtype=torch.float16)
    vae.encode(imgs).latent_dist.mean          vae.decode(latents).sample

plus the fused wrappers ``encode_imgs(imgs)`` / ``decode_latents(latents)`` that fold the reference's surrounding
arithmetic (``2*imgs-1``, ``*0.18215``, ``/0.18215``, ``(x/2+0.5).clamp(0,1)``) into the staging kernels.

Every convolution / linear is tcl_igemm (tcgen05 implicit GEMM), norms are tcl_groupnorm; ResnetBlock2D fuses the
1x1 shortcut into conv2's K loop like the UNet.  The mid-block attention has ONE head of width 512 — outside
tcl_attention's TMEM budget (d_pad <= 192) — so it runs as three launches: scores = Q K^T (tcl_igemm, N = tokens),
tcl_softmax_rows, O = P V (tcl_igemm with V^T as the weight); V^T is produced directly by an igemm with swapped roles
and V's bias is folded into the output projection's bias (softmax rows sum to one).
Activations are NHWC 16-bit; 3 / 4 / 8-channel tensors are padded to 64 channels.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn

from . import ops
from ._lib import TclError, check, dtype_code, latent_code, lib, require_cuda, stream_ptr
from .weights import pack_conv3x3


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _Conv:
    """3x3 (or 1x1) conv weights repacked for tcl_igemm; in-channels padded to a multiple of 64, out-channels to 64."""

    def __init__(self, w, b, dev, dt, out_pad: int = 64):
        co, ci, kh, kw = w.shape
        cip = (ci + 63) // 64 * 64
        cop = (co + out_pad - 1) // out_pad * out_pad
        wp = torch.zeros(cop, cip, kh, kw, dtype=torch.float32)
        wp[:co, :ci] = w.detach().float()
        self.taps = kh * kw
        self.w = (pack_conv3x3(wp) if self.taps == 9 else wp.reshape(cop, cip)).to(device=dev, dtype=dt).contiguous()
        bp = torch.zeros(cop)
        bp[:co] = b.detach().float()
        self.b = bp.to(dev).contiguous()
        self.cout, self.cout_pad, self.cin_pad = co, cop, cip


class _Resnet(nn.Module):
    def __init__(self, sd, prefix, dev, dt, groups=32):
        super().__init__()
        g = lambda k: sd[prefix + k]
        self.groups = groups
        self.n1_w, self.n1_b = _f32(g("norm1.weight"), dev), _f32(g("norm1.bias"), dev)
        self.n2_w, self.n2_b = _f32(g("norm2.weight"), dev), _f32(g("norm2.bias"), dev)
        self.conv1 = _Conv(g("conv1.weight"), g("conv1.bias"), dev, dt)
        w2 = pack_conv3x3(g("conv2.weight").detach().float())
        b2 = g("conv2.bias").detach().float()
        self.has_shortcut = (prefix + "conv_shortcut.weight") in sd
        if self.has_shortcut:
            ws = g("conv_shortcut.weight").detach().float()
            w2 = torch.cat([w2, ws.reshape(ws.shape[0], ws.shape[1])], dim=1)
            b2 = b2 + g("conv_shortcut.bias").detach().float()
        self.w2 = w2.to(device=dev, dtype=dt).contiguous()
        self.b2 = b2.to(dev).contiguous()

    def forward(self, x):
        n, h, w, _ = x.shape
        hn = ops.groupnorm(x, self.n1_w, self.n1_b, self.groups, 1e-6, True)
        h1 = ops.igemm([(hn, 9, 1)], self.conv1.w, (n, h, w), bias=self.conv1.b)
        hn2 = ops.groupnorm(h1, self.n2_w, self.n2_b, self.groups, 1e-6, True)
        if self.has_shortcut:
            return ops.igemm([(hn2, 9, 1), (x, 1, 1)], self.w2, (n, h, w), bias=self.b2)
        return ops.igemm([(hn2, 9, 1)], self.w2, (n, h, w), bias=self.b2, residual=x)


class _Attention(nn.Module):
    def __init__(self, sd, prefix, dev, dt, groups=32):
        super().__init__()
        g = lambda k: sd[prefix + k]
        cvt = lambda t: t.detach().to(device=dev, dtype=dt).contiguous()
        self.groups, self.dt = groups, dt
        self.gn_w, self.gn_b = _f32(g("group_norm.weight"), dev), _f32(g("group_norm.bias"), dev)
        self.C = g("to_q.weight").shape[0]
        self.wq, self.bq = cvt(g("to_q.weight")), _f32(g("to_q.bias"), dev)
        self.wk, self.bk = cvt(g("to_k.weight")), _f32(g("to_k.bias"), dev)
        self.wv = cvt(g("to_v.weight")).view(1, 1, self.C, self.C)
        wo = g("to_out.0.weight").detach().float()
        self.wo = cvt(wo)
        # softmax rows sum to 1  =>  P (V + 1 bv^T) Wo^T + bo = (P V) Wo^T + (Wo bv + bo)
        self.bo = (wo @ g("to_v.bias").detach().float() + g("to_out.0.bias").detach().float()).to(dev).contiguous()

    def forward(self, x):
        n, h, w, C = x.shape
        T = h * w
        Tp = (T + 63) // 64 * 64
        t = ops.groupnorm(x, self.gn_w, self.gn_b, self.groups, 1e-6, False)
        out = torch.empty_like(x)
        kbuf = torch.zeros((Tp, C), device=x.device, dtype=self.dt)
        tpad = torch.zeros((Tp, C), device=x.device, dtype=self.dt) if Tp != T else None
        scores = torch.empty((T, Tp), device=x.device, dtype=self.dt)
        for i in range(n):                      # one image at a time: scores are T x T
            ti = t[i].view(T, C)
            q = ops.igemm([(ti.view(1, 1, T, C), 1, 1)], self.wq, (1, 1, T), bias=self.bq, out_scale=1.0 / math.sqrt(C))[0, 0]
            ops.linear(ti, self.wk, bias=self.bk, out=kbuf[:T])
            if tpad is not None:
                tpad[:T].copy_(ti)
                tw = tpad
            else:
                tw = ti
            vt = ops.igemm([(self.wv, 1, 1)], tw, (1, 1, C))[0, 0]                     # V^T [C, Tp]  (bias folded into bo)
            ops.igemm([(q.view(1, 1, T, C), 1, 1)], kbuf, (1, 1, T), out=scores.view(1, 1, T, Tp))
            check(lib.tcl_softmax_rows(dtype_code(self.dt), scores.data_ptr(), T, T, Tp, stream_ptr()), "tcl_softmax_rows")
            o = ops.igemm([(scores.view(1, 1, T, Tp), 1, 1)], vt, (1, 1, T))[0, 0]     # [T, C]
            ops.linear(o, self.wo, bias=self.bo, residual=x[i].view(T, C), out=out[i].view(T, C))
        return out


class _Stack(nn.Module):
    pass


class AutoencoderKLB200(nn.Module):
    scaling_factor = 0.18215

    def __init__(self, sd: Dict[str, torch.Tensor], device="cuda", dtype=torch.float16, block_out_channels=(128, 256, 512, 512),
                 latent_channels: int = 4):
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise TclError("AutoencoderKLB200 runs on CUDA only (no CPU path)")
        boc = tuple(block_out_channels)
        if any(c % 64 for c in boc):
            raise TclError("AutoencoderKLB200: block_out_channels must be multiples of 64")
        self.dev, self.dt, self.boc, self.latent = dev, dtype, boc, latent_channels
        R = lambda p: _Resnet(sd, p, dev, dtype)
        C_ = lambda p: _Conv(sd[p + ".weight"], sd[p + ".bias"], dev, dtype)
        # ---- encoder
        self.e_conv_in = C_("encoder.conv_in")
        self.e_down = []
        for i in range(len(boc)):
            blk = _Stack()
            blk.resnets = [R(f"encoder.down_blocks.{i}.resnets.{j}.") for j in range(2)]
            blk.down = C_(f"encoder.down_blocks.{i}.downsamplers.0.conv") if i < len(boc) - 1 else None
            self.e_down.append(blk)
        self.e_mid = [R("encoder.mid_block.resnets.0."), _Attention(sd, "encoder.mid_block.attentions.0.", dev, dtype),
                      R("encoder.mid_block.resnets.1.")]
        self.e_no_w, self.e_no_b = _f32(sd["encoder.conv_norm_out.weight"], dev), _f32(sd["encoder.conv_norm_out.bias"], dev)
        self.e_conv_out = C_("encoder.conv_out")
        self.quant = C_("quant_conv")
        # ---- decoder
        self.post_quant = C_("post_quant_conv")
        self.d_conv_in = C_("decoder.conv_in")
        self.d_mid = [R("decoder.mid_block.resnets.0."), _Attention(sd, "decoder.mid_block.attentions.0.", dev, dtype),
                      R("decoder.mid_block.resnets.1.")]
        self.d_up = []
        for i in range(len(boc)):
            blk = _Stack()
            blk.resnets = [R(f"decoder.up_blocks.{i}.resnets.{j}.") for j in range(3)]
            blk.up = C_(f"decoder.up_blocks.{i}.upsamplers.0.conv") if i < len(boc) - 1 else None
            self.d_up.append(blk)
        self.d_no_w, self.d_no_b = _f32(sd["decoder.conv_norm_out.weight"], dev), _f32(sd["decoder.conv_norm_out.bias"], dev)
        # full-resolution output: keep only 8 channels (3 real) instead of a 64-wide NHWC image
        self.d_conv_out = _Conv(sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], dev, dtype, out_pad=8)

    # diffusers-style attributes the reference reads (generate.py:574)
    @property
    def device(self):
        return self.dev

    @property
    def dtype(self):
        return self.dt

    # ---- staging --------------------------------------------------------------------------------------
    def _to_nhwc(self, x: torch.Tensor, scale: float, shift: float) -> torch.Tensor:
        require_cuda(x)
        x = x.contiguous()
        B, C, H, W = x.shape
        out = torch.empty((B, H, W, 64), device=x.device, dtype=self.dt)
        check(lib.tcl_image_to_nhwc(dtype_code(self.dt), latent_code(x.dtype), x.data_ptr(), B, C, H, W, 64, scale, shift,
                                    out.data_ptr(), stream_ptr()), "tcl_image_to_nhwc")
        return out

    def _to_nchw(self, y: torch.Tensor, C: int, out_dtype, scale: float, shift: float, clamp: bool) -> torch.Tensor:
        B, H, W, pitch = y.shape
        out = torch.empty((B, C, H, W), device=y.device, dtype=out_dtype)
        check(lib.tcl_nhwc_to_image(dtype_code(self.dt), y.data_ptr(), B, C, H, W, pitch, scale, shift, int(clamp), 0.0, 1.0,
                                    latent_code(out_dtype), out.data_ptr(), stream_ptr()), "tcl_nhwc_to_image")
        return out

    # ---- networks -------------------------------------------------------------------------------------
    def _conv(self, c: _Conv, x, stride=1, no_lead_pad=False, out_hw=None):
        n, h, w, _ = x.shape
        oh, ow = (h, w) if out_hw is None else out_hw
        return ops.igemm([(x, c.taps, stride, no_lead_pad)], c.w, (n, oh, ow), bias=c.b)

    def _encode_nhwc(self, x):
        """x: NHWC [B,H,W,64] (3 real channels) -> moments NHWC [B,h,w,64] (2*latent real channels)."""
        h = self._conv(self.e_conv_in, x)
        for blk in self.e_down:
            for r in blk.resnets:
                h = r(h)
            if blk.down is not None:
                n, hh, ww, _ = h.shape
                # F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv without padding: out = floor((H + 1 - 3) / 2) + 1
                h = self._conv(blk.down, h, stride=2, no_lead_pad=True, out_hw=((hh - 2) // 2 + 1, (ww - 2) // 2 + 1))
        for m in self.e_mid:
            h = m(h)
        h = ops.groupnorm(h, self.e_no_w, self.e_no_b, 32, 1e-6, True)
        h = self._conv(self.e_conv_out, h)                 # [.., 64] (8 real)
        return self._conv(self.quant, h)                   # 1x1

    def _decode_nhwc(self, z):
        """z: NHWC [B,h,w,64] (latent real channels) -> image NHWC [B,H,W,64] (3 real channels)."""
        h = self._conv(self.post_quant, z)
        h = self._conv(self.d_conv_in, h)
        for m in self.d_mid:
            h = m(h)
        for blk in self.d_up:
            for r in blk.resnets:
                h = r(h)
            if blk.up is not None:
                n, hh, ww, _ = h.shape
                h = self._conv(blk.up, ops.upsample_nearest(h, 2 * hh, 2 * ww))
        h = ops.groupnorm(h, self.d_no_w, self.d_no_b, 32, 1e-6, True)
        return self._conv(self.d_conv_out, h)

    # ---- diffusers-compatible operators ----------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor):
        mom = self._encode_nhwc(self._to_nhwc(x, 1.0, 0.0))
        m = self._to_nchw(mom, 2 * self.latent, x.dtype if x.dtype != torch.float32 else self.dt, 1.0, 0.0, False)
        mean, logvar = m.chunk(2, dim=1)
        dist = type("DiagonalGaussian", (), {"mean": mean, "logvar": logvar, "mode": lambda self_: mean})()
        return type("EncoderOutput", (), {"latent_dist": dist})()

    @torch.no_grad()
    def decode(self, z: torch.Tensor):
        img = self._decode_nhwc(self._to_nhwc(z, 1.0, 0.0))
        out = self._to_nchw(img, 3, z.dtype if z.dtype != torch.float32 else self.dt, 1.0, 0.0, False)
        return type("DecoderOutput", (), {"sample": out})()

    # ---- the reference's wrappers, fused (generate_utils.py:140-172) --------------------------------------
    @torch.no_grad()
    def encode_imgs(self, imgs: torch.Tensor, batch_size: int = 2) -> torch.Tensor:
        """imgs [N,3,H,W] in [0,1] -> latents [N,4,H/8,W/8] = posterior.mean * 0.18215."""
        outs = []
        for b in imgs.split(batch_size, dim=0):
            mom = self._encode_nhwc(self._to_nhwc(b, 2.0, -1.0))
            outs.append(self._to_nchw(mom, self.latent, self.dt, self.scaling_factor, 0.0, False))
        return torch.cat(outs)

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor, batch_size: int = 2) -> torch.Tensor:
        """latents [N,4,h,w] -> images [N,3,8h,8w] = (decode(latents / 0.18215) / 2 + 0.5).clamp(0, 1)."""
        outs = []
        for b in latents.split(batch_size, dim=0):
            img = self._decode_nhwc(self._to_nhwc(b, 1.0 / self.scaling_factor, 0.0))
            outs.append(self._to_nchw(img, 3, self.dt, 0.5, 0.5, True))
        return torch.cat(outs)
