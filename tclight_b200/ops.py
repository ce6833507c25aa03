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

    srcs      : [(NHWC tensor [n,h,w,c], taps (1|9), stride (1|2)), ...] — K segments in order
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
    for i, (t, taps, stride) in enumerate(srcs):
        require_cuda(t)
        if t.dtype != dt or t.dim() != 4:
            raise TclError(f"igemm source {i}: need 4-D NHWC {dt}, got {tuple(t.shape)} {t.dtype}")
        s = d.src[i]
        s.ptr = t.data_ptr()
        s.n, s.h, s.w, s.c = t.shape
        s.pitch = _pixel_pitch(t)
        s.taps = taps
        s.stride = stride
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
    check(lib.tcl_attention(C.byref(a), stream_ptr()), "tcl_attention")
    return out
