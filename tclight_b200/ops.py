"""Thin torch-tensor wrappers over the C ABI (include/tclight.h).

Tensors carry device memory only; all arithmetic happens inside libtclight.so.  Every wrapper
validates shapes/dtypes, fills the plain-C descriptor, and raises ``TclError`` on failure.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import TclError, check, dtype_code, lib, require_cuda, stream_ptr


# ---------------------------------------------------------------------------------------------
# optional per-kernel profiling (bench.py): CUDA events around each launch on the launching
# stream + algorithmic FLOP counts.  Off by default (zero overhead).
# ---------------------------------------------------------------------------------------------
_PROFILE = None


def profile_start():
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """-> {name: {flops, ms, launches}} (synchronises)."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    out = {}
    if not prof:
        return out
    torch.cuda.synchronize()
    for name, recs in prof.items():
        out[name] = {"flops": float(sum(r[2] for r in recs)), "ms": float(sum(r[0].elapsed_time(r[1]) for r in recs)),
                     "launches": len(recs)}
    return out


class _Prof:
    def __init__(self, name, flops):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if _PROFILE is not None:
            self.e.record()
            _PROFILE.setdefault(self.name, []).append((self.s, self.e, self.flops))
        return False


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _pixel_pitch(t: torch.Tensor) -> int:
    """NHWC tensor (possibly a channel-slice view): elements between consecutive pixels."""
    n, h, w, c = t.shape
    if t.stride(3) != 1:
        raise TclError("igemm source must have unit channel stride")
    pitch = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
    if w > 1 and h > 1 and t.stride(1) != w * pitch:
        raise TclError("igemm source rows must be dense in w")
    if n > 1 and t.stride(0) != h * w * pitch and h * w > 1:
        raise TclError("igemm source images must be dense")
    return pitch


def igemm(
    srcs: Sequence[Tuple[torch.Tensor, int, int]],
    weight: torch.Tensor,
    out_grid: Tuple[int, int, int],
    bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    mode: int = L.TCL_EPI_NHWC,
    out: Optional[torch.Tensor] = None,
    out_scale: float = 1.0,
    heads: Optional[dict] = None,
) -> Optional[torch.Tensor]:
    """out[pixel, n] = sum_k A[pixel, k] W[n, k]  (tcgen05 implicit GEMM).

    srcs      : [(NHWC tensor [n,h,w,c], taps (1|9), stride (1|2)[, no_lead_pad]), ...] — K segments in order
    weight    : [N, K] 16-bit, K = sum(taps*c)
    out_grid  : (n_img, out_h, out_w)
    heads     : for TCL_EPI_HEADS: dict(sec=[(tensor, is_vt), ...], heads=, d=, d_pad=,
                tok_per_batch=, tok_pitch=)
    """
    dt = weight.dtype
    d = L.IgemmDesc()
    d.dtype = dtype_code(dt)
    d.num_src = len(srcs)
    ktot = 0
    for i, src in enumerate(srcs):
        t, taps, stride = src[:3]
        require_cuda(t)
        if t.dtype != dt or t.dim() != 4:
            raise TclError(f"igemm source {i}: need 4-D NHWC {dt}, got {tuple(t.shape)} {t.dtype}")
        s = d.src[i]
        s.ptr = t.data_ptr()
        s.n, s.h, s.w, s.c = t.shape
        s.pitch = _pixel_pitch(t)
        s.taps = taps
        s.stride = stride
        s.no_lead_pad = int(len(src) > 3 and bool(src[3]))
        ktot += taps * t.shape[3]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != ktot or not weight.is_contiguous():
        raise TclError(f"igemm weight must be contiguous [N, {ktot}], got {tuple(weight.shape)}")
    require_cuda(weight, bias, residual, out)
    d.n_img, d.out_h, d.out_w = out_grid
    d.N = N
    d.K = ktot
    d.weight = weight.data_ptr()
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise TclError("igemm bias must be fp32 [N]")
        d.bias = bias.data_ptr()
    d.mode = mode
    d.out_scale = out_scale
    n_img, oh, ow = out_grid
    if mode == L.TCL_EPI_HEADS:
        assert heads is not None
        for i, (t, is_vt) in enumerate(heads["sec"]):
            require_cuda(t)
            d.sec_ptr[i] = t.data_ptr()
            d.sec_vt[i] = int(is_vt)
        d.sec_cols = heads["heads"] * heads["d"]
        d.heads = heads["heads"]
        d.d = heads["d"]
        d.d_pad = heads["d_pad"]
        d.tok_per_batch = heads["tok_per_batch"]
        d.tok_pitch = heads["tok_pitch"]
        ret = None
    else:
        n_out = N // 2 if mode == L.TCL_EPI_GEGLU else N
        if out is None:
            out = torch.empty((n_img, oh, ow, n_out), device=weight.device, dtype=dt)
        if out.dtype != dt or out.shape[-1] != n_out or out.stride(-1) != 1:
            raise TclError("igemm out tensor mismatch")
        d.out = out.data_ptr()
        d.out_pitch = out.stride(-2) if out.dim() >= 2 else n_out
        if residual is not None:
            if residual.dtype != dt or residual.shape[-1] != N:
                raise TclError("igemm residual mismatch")
            d.residual = residual.data_ptr()
            d.res_pitch = residual.stride(-2)
        ret = out
    with _Prof("igemm_conv3x3" if any(sr[1] == 9 for sr in srcs) else "igemm_linear", 2.0 * n_img * oh * ow * N * ktot):
        check(lib.tcl_igemm(C.byref(d), stream_ptr()), "tcl_igemm")
    return ret


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None, residual=None, out=None, mode=L.TCL_EPI_NHWC):
    """x [M, K] @ weight[N, K]^T (+bias)(+residual) -> [M, N] (or [M, N/2] for GEGLU)."""
    M, K = x.shape
    src = x.view(1, 1, M, K)
    n_out = weight.shape[0] // 2 if mode == L.TCL_EPI_GEGLU else weight.shape[0]
    if out is None:
        out = torch.empty((M, n_out), device=x.device, dtype=x.dtype)
    res4 = None if residual is None else residual.view(1, 1, M, -1)
    igemm([(src, 1, 1)], weight, (1, 1, M), bias=bias, residual=res4, mode=mode, out=out.view(1, 1, M, n_out))
    return out


def head_pad(d: int) -> int:
    """Head dim padded to a multiple of the 64-element swizzle chunk (40->64, 80->128, 160->192)."""
    dp = (d + 63) // 64 * 64
    if dp > 192:
        raise TclError(f"head_dim {d} not supported (max 192)")
    return dp


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, tq: int, tk: int, d: int,
              kv_batch_div: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v on head-split operands (see tcl_attention in tclight.h).

    q [B, H, tq_pitch, d_pad], k [Bkv, H, tk_pitch, d_pad], vt [Bkv, H, d_pad, tk_pitch]
    -> out [B, tq, H*d]
    """
    require_cuda(q, k, vt)
    B, H, tq_pitch, d_pad = q.shape
    Bkv, _, tk_pitch, _ = k.shape
    if vt.shape != (Bkv, H, d_pad, tk_pitch) or not (q.is_contiguous() and k.is_contiguous() and vt.is_contiguous()):
        raise TclError("attention operands must be contiguous head-split tensors")
    if Bkv * kv_batch_div != B:
        raise TclError("attention: kv batch mismatch")
    if out is None:
        out = torch.empty((B, tq, H * d), device=q.device, dtype=q.dtype)
    a = L.AttnDesc()
    a.dtype = dtype_code(q.dtype)
    a.batch, a.heads, a.tq, a.tk, a.d, a.d_pad = B, H, tq, tk, d, d_pad
    a.kv_batch_div = kv_batch_div
    a.tq_pitch, a.tk_pitch = tq_pitch, tk_pitch
    a.q, a.k, a.vt, a.out = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr()
    ws = _attn_ws.get(q.device)
    if ws is None:
        ws = _attn_ws[q.device] = torch.empty(L.lib.tcl_attention_workspace_bytes(), device=q.device, dtype=torch.uint8)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    with _Prof(f"attention_d{d_pad}", 4.0 * B * H * tq * tk * d):
        check(lib.tcl_attention(C.byref(a), stream_ptr()), "tcl_attention")
    return out


_gn_ws = {}
_attn_ws = {}        # per device: scratch of the KV-split attention tail (tcl_attention_workspace_bytes)


def _stats_ws(dev, n):
    key = (dev, n >= 0)
    ws = _gn_ws.get(dev)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 4096), device=dev, dtype=torch.float32)
        _gn_ws[dev] = ws
    return ws


def groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
              x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over NHWC [n,h,w,c1] (optionally channel-concatenated with x2 [n,h,w,c2])."""
    require_cuda(x1, x2, gamma, beta)
    n, h, w, c1 = x1.shape
    c2 = 0 if x2 is None else x2.shape[3]
    if not x1.is_contiguous() or (x2 is not None and not x2.is_contiguous()):
        raise TclError("groupnorm inputs must be contiguous NHWC")
    if gamma.dtype != torch.float32 or gamma.numel() != c1 + c2:
        raise TclError("groupnorm affine must be fp32 [C]")
    if out is None:
        out = torch.empty((n, h, w, c1 + c2), device=x1.device, dtype=x1.dtype)
    ws = _stats_ws(x1.device, 2 * groups * n)
    check(lib.tcl_groupnorm(dtype_code(x1.dtype), x1.data_ptr(), c1, 0 if x2 is None else x2.data_ptr(), c2, n, h * w,
                            groups, gamma.data_ptr(), beta.data_ptr(), eps, int(silu), ws.data_ptr(), out.data_ptr(),
                            stream_ptr()), "tcl_groupnorm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, gamma, beta)
    C_ = x.shape[-1]
    if not x.is_contiguous():
        raise TclError("layernorm input must be contiguous")
    rows = x.numel() // C_
    if out is None:
        out = torch.empty_like(x)
    check(lib.tcl_layernorm(dtype_code(x.dtype), x.data_ptr(), rows, C_, gamma.data_ptr(), beta.data_ptr(), eps,
                            out.data_ptr(), stream_ptr()), "tcl_layernorm")
    return out


def upsample_nearest(x: torch.Tensor, oh: int, ow: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x)
    n, h, w, c = x.shape
    if not x.is_contiguous():
        raise TclError("upsample input must be contiguous NHWC")
    if out is None:
        out = torch.empty((n, oh, ow, c), device=x.device, dtype=x.dtype)
    check(lib.tcl_upsample_nearest(x.data_ptr(), n, h, w, c, oh, ow, out.data_ptr(), stream_ptr()),
          "tcl_upsample_nearest")
    return out


def _strides4(t: torch.Tensor):
    """Element strides {image, channel, row, col} of a 4-D [F, 4, H, W] latent view."""
    arr = (C.c_longlong * 4)(*[int(s) for s in t.stride()])
    return arr


def stage_latent(x: torch.Tensor, cond: Optional[torch.Tensor], act_dtype, out: Optional[torch.Tensor] = None,
                 duplicate: bool = True):
    """x, cond: [F, 4, H, W] (any strides) -> NHWC [2F, H, W, 64] act_dtype (CFG halves duplicated),
    or [F, H, W, 64] with duplicate=False."""
    require_cuda(x, cond)
    F_, c, H, W = x.shape
    if c != 4 or (cond is not None and (cond.shape != x.shape or cond.dtype != x.dtype)):
        raise TclError("stage_latent expects matching [F,4,H,W] latents")
    if out is None:
        out = torch.empty(((2 if duplicate else 1) * F_, H, W, 64), device=x.device, dtype=act_dtype)
    check(lib.tcl_stage_latent(dtype_code(act_dtype), L.latent_code(x.dtype), x.data_ptr(), _strides4(x),
                               0 if cond is None else cond.data_ptr(), None if cond is None else _strides4(cond),
                               F_, H, W, int(duplicate), out.data_ptr(), stream_ptr()), "tcl_stage_latent")
    return out


def cfg_store(eps: torch.Tensor, guidance_scale: float, out_view: torch.Tensor) -> None:
    """eps NHWC [2F, H, W, pitch>=4] -> out_view[F,4,H,W] = uncond + g (cond - uncond)."""
    require_cuda(eps, out_view)
    F2, H, W, pitch = eps.shape
    F_ = F2 // 2
    if out_view.shape != (F_, 4, H, W) or not eps.is_contiguous():
        raise TclError("cfg_store shape mismatch")
    check(lib.tcl_cfg_store(dtype_code(eps.dtype), L.latent_code(out_view.dtype), eps.data_ptr(), pitch,
                            float(guidance_scale), F_, H, W, out_view.data_ptr(), _strides4(out_view), stream_ptr()),
          "tcl_cfg_store")


# ---------------------------------------------------------------------------------------------
# VidToMe primitives
# ---------------------------------------------------------------------------------------------
def normalize_split(x0: torch.Tensor, x1: Optional[torch.Tensor], d0: int, d1: int):
    """tokens = [x0 | x1] ([B, n0, C], [B, n1, C]); dst = tokens[:, d0:d1]; returns normalised
    (a [B, n_src, C], b [B, n_dst, C])  (merge.py:84-85)."""
    require_cuda(x0, x1)
    B, n0, C_ = x0.shape
    n1 = 0 if x1 is None else x1.shape[1]
    if not x0.is_contiguous() or (x1 is not None and not x1.is_contiguous()):
        raise TclError("normalize_split inputs must be contiguous")
    n_dst = d1 - d0
    n_src = n0 + n1 - n_dst
    a = torch.empty((B, n_src, C_), device=x0.device, dtype=x0.dtype)
    b = torch.empty((B, n_dst, C_), device=x0.device, dtype=x0.dtype)
    check(lib.tcl_vidtome_normalize_split(dtype_code(x0.dtype), x0.data_ptr(), n0, 0 if x1 is None else x1.data_ptr(),
                                          n1, B, C_, d0, d1, a.data_ptr(), b.data_ptr(), stream_ptr()),
          "tcl_vidtome_normalize_split")
    return a, b


def vidtome_match(a: torch.Tensor, b: torch.Tensor, align_batch: bool):
    """node_max (fp32 holding 16-bit-rounded scores), node_idx (int64) — merge.py:87-97."""
    require_cuda(a, b)
    B, n_src, C_ = a.shape
    n_dst = b.shape[1]
    shape = (n_src,) if align_batch else (B, n_src)
    node_max = torch.empty(shape, device=a.device, dtype=torch.float32)
    node_idx = torch.empty(shape, device=a.device, dtype=torch.int64)
    wsb = lib.tcl_vidtome_match_workspace_bytes(B, n_src)
    ws = torch.empty(wsb, device=a.device, dtype=torch.uint8)
    with _Prof("vidtome_match", 2.0 * B * n_src * n_dst * C_):
        check(lib.tcl_vidtome_match(dtype_code(a.dtype), a.data_ptr(), b.data_ptr(), B, n_src, n_dst, C_, int(align_batch),
                                    node_max.data_ptr(), node_idx.data_ptr(), ws.data_ptr(), wsb, stream_ptr()),
              "tcl_vidtome_match")
    return node_max, node_idx


def vidtome_plan(edge: torch.Tensor, node_idx: torch.Tensor, n_src: int, n_dst: int, r: int, d0: int):
    """-> (merge_map int32 [n_src-r+n_dst], unmerge_map int32 [n_src+n_dst])."""
    require_cuda(edge, node_idx)
    mm = torch.empty(n_src - r + n_dst, device=edge.device, dtype=torch.int32)
    um = torch.empty(n_src + n_dst, device=edge.device, dtype=torch.int32)
    check(lib.tcl_vidtome_plan(edge.data_ptr(), node_idx.data_ptr(), n_src, n_dst, r, d0, mm.data_ptr(), um.data_ptr(),
                               stream_ptr()), "tcl_vidtome_plan")
    return mm, um


def gather_rows(x0: torch.Tensor, x1: Optional[torch.Tensor], idx_map: torch.Tensor, add: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b, i] = [x0 | x1][b, map[i]] (+ add[b, i]);  map int32 [n_out] (shared) or [B, n_out]."""
    require_cuda(x0, x1, idx_map, add)
    B, n0, C_ = x0.shape
    n1 = 0 if x1 is None else x1.shape[1]
    per_batch = idx_map.dim() == 2
    n_out = idx_map.shape[-1]
    if idx_map.dtype != torch.int32 or not idx_map.is_contiguous():
        raise TclError("gather map must be contiguous int32")
    if not x0.is_contiguous() or (x1 is not None and not x1.is_contiguous()) or (add is not None and not add.is_contiguous()):
        raise TclError("gather_rows inputs must be contiguous")
    if out is None:
        out = torch.empty((B, n_out, C_), device=x0.device, dtype=x0.dtype)
    check(lib.tcl_gather_rows(dtype_code(x0.dtype), x0.data_ptr(), n0, 0 if x1 is None else x1.data_ptr(), n1,
                              idx_map.data_ptr(), int(per_batch), n_out, B, C_, 0 if add is None else add.data_ptr(),
                              out.data_ptr(), stream_ptr()), "tcl_gather_rows")
    return out


def gemv(W: torch.Tensor, x: torch.Tensor, bias: Optional[torch.Tensor], silu_in: bool, round16: bool = True,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = W f(x) + b for one fp32 vector x [K]; W [N, K] 16-bit; returns fp32 [N]."""
    require_cuda(W, x, bias)
    N, K = W.shape
    if x.dtype != torch.float32 or x.numel() != K or not W.is_contiguous():
        raise TclError("gemv: x must be fp32 [K]")
    if out is None:
        out = torch.empty(N, device=W.device, dtype=torch.float32)
    check(lib.tcl_gemv(dtype_code(W.dtype), W.data_ptr(), x.data_ptr(), 0 if bias is None else bias.data_ptr(), N, K,
                       int(silu_in), int(round16), out.data_ptr(), stream_ptr()), "tcl_gemv")
    return out


# ---------------------------------------------------------------------------------------------
# sampler tail
# ---------------------------------------------------------------------------------------------
def adain_blend(noises_t: torch.Tensor, noises: torch.Tensor, alpha: float) -> None:
    """In place: noises_t <- AdaIN(noises_t, noises); noises <- sqrt(a) noises_t + sqrt(1-a) noises."""
    require_cuda(noises_t, noises)
    if noises_t.shape != noises.shape or noises_t.dtype != noises.dtype or noises.dim() != 4:
        raise TclError("adain_blend: shape/dtype mismatch")
    if not (noises_t.is_contiguous() and noises.is_contiguous()):
        raise TclError("adain_blend: latents must be contiguous")
    N, Cc, h, w = noises.shape
    check(lib.tcl_adain_blend(L.latent_code(noises.dtype), noises_t.data_ptr(), noises.data_ptr(), N * Cc, h * w,
                              float(alpha), stream_ptr()), "tcl_adain_blend")


def scale_inplace(x: torch.Tensor, s: float) -> None:
    require_cuda(x)
    if not x.is_contiguous():
        raise TclError("scale_inplace: tensor must be contiguous")
    check(lib.tcl_scale_inplace(L.latent_code(x.dtype), x.data_ptr(), x.numel(), float(s), stream_ptr()),
          "tcl_scale_inplace")


def dpm_step(eps, x, x0_prev, z, coef: dict):
    """One DPM-Solver++ SDE step; returns (x0, x_next) in the latent dtype."""
    require_cuda(eps, x, x0_prev, z)
    if z.dtype != torch.float32 or not (eps.is_contiguous() and x.is_contiguous() and z.is_contiguous()):
        raise TclError("dpm_step: z must be contiguous fp32, latents contiguous")
    x0 = torch.empty_like(eps)
    xn = torch.empty_like(eps)
    check(lib.tcl_dpm_step(L.latent_code(eps.dtype), eps.data_ptr(), x.data_ptr(), 0 if x0_prev is None else x0_prev.data_ptr(),
                           z.data_ptr(), x0.data_ptr(), xn.data_ptr(), eps.numel(), coef["sigma_c_hat"], coef["alpha_c_hat"],
                           coef["A"], coef["B"], coef["Cn"], coef["inv_r0"], int(coef["second_order"]), stream_ptr()),
          "tcl_dpm_step")
    return x0, xn
