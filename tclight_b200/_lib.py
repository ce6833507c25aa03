"""ctypes binding of libtclight.so (the C ABI declared in include/tclight.h).

The library is the product: there is no Python/PyTorch fallback.  Importing this module
without a built ``libtclight.so`` raises, and every wrapper raises ``TclError`` when a call
returns a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtclight.so")

TCL_DTYPE_FP16 = 0
TCL_DTYPE_BF16 = 1
TCL_EPI_NHWC = 0
TCL_EPI_GEGLU = 1
TCL_EPI_HEADS = 2
TCL_IGEMM_MAX_SRC = 4


class TclError(RuntimeError):
    pass


class IgemmSrc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("n", C.c_int64),
        ("h", C.c_int64),
        ("w", C.c_int64),
        ("c", C.c_int64),
        ("pitch", C.c_int64),
        ("taps", C.c_int32),
        ("stride", C.c_int32),
        ("no_lead_pad", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class IgemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("num_src", C.c_int32),
        ("src", IgemmSrc * TCL_IGEMM_MAX_SRC),
        ("n_img", C.c_int32),
        ("out_h", C.c_int32),
        ("out_w", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int64),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("mode", C.c_int32),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int64),
        ("residual", C.c_void_p),
        ("res_pitch", C.c_int64),
        ("out_scale", C.c_float),
        ("sec_ptr", C.c_void_p * 3),
        ("sec_vt", C.c_int32 * 3),
        ("sec_cols", C.c_int32),
        ("heads", C.c_int32),
        ("d", C.c_int32),
        ("d_pad", C.c_int32),
        ("tok_per_batch", C.c_int64),
        ("tok_pitch", C.c_int64),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tclight_b200/csrc`). tclight_b200 has no CPU/PyTorch fallback."
        )
    return C.CDLL(LIB_PATH)


lib = _load()
lib.tcl_last_error.restype = C.c_char_p
lib.tcl_version.restype = C.c_int
lib.tcl_launch_count.restype = C.c_longlong
lib.tcl_launch_count_reset.restype = None


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.tcl_last_error().decode("utf-8", "replace")
        raise TclError(f"{what} failed (status {rc}): {msg}")


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t) -> int:
    import torch

    if t == torch.float16:
        return TCL_DTYPE_FP16
    if t == torch.bfloat16:
        return TCL_DTYPE_BF16
    raise TclError(f"unsupported activation dtype {t}; tclight kernels take fp16 or bf16")


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TclError("tclight_b200 ops need CUDA tensors (no CPU path exists)")


lib.tcl_igemm.argtypes = [C.POINTER(IgemmDesc), C.c_void_p]
lib.tcl_igemm.restype = C.c_int


class AttnDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("batch", C.c_int32),
        ("heads", C.c_int32),
        ("tq", C.c_int32),
        ("tk", C.c_int32),
        ("d", C.c_int32),
        ("d_pad", C.c_int32),
        ("kv_batch_div", C.c_int32),
        ("tq_pitch", C.c_int64),
        ("tk_pitch", C.c_int64),
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("vt", C.c_void_p),
        ("out", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


lib.tcl_attention.argtypes = [C.POINTER(AttnDesc), C.c_void_p]
lib.tcl_attention.restype = C.c_int
lib.tcl_attention_workspace_bytes.argtypes = []
lib.tcl_attention_workspace_bytes.restype = C.c_size_t

TCL_LATENT_FP32 = 0
TCL_LATENT_FP16 = 1
TCL_LATENT_BF16 = 2

lib.tcl_groupnorm.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tcl_groupnorm.restype = C.c_int
lib.tcl_layernorm.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                              C.c_void_p, C.c_void_p]
lib.tcl_layernorm.restype = C.c_int
lib.tcl_upsample_nearest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]
lib.tcl_upsample_nearest.restype = C.c_int
lib.tcl_stage_latent.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p,
                                 C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_stage_latent.restype = C.c_int
lib.tcl_cfg_store.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p]
lib.tcl_cfg_store.restype = C.c_int


def latent_code(t) -> int:
    import torch

    if t == torch.float32:
        return TCL_LATENT_FP32
    if t == torch.float16:
        return TCL_LATENT_FP16
    if t == torch.bfloat16:
        return TCL_LATENT_BF16
    raise TclError(f"unsupported latent dtype {t}")

lib.tcl_vidtome_match_workspace_bytes.argtypes = [C.c_int, C.c_int]
lib.tcl_vidtome_match_workspace_bytes.restype = C.c_size_t
lib.tcl_vidtome_normalize_split.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                            C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tcl_vidtome_normalize_split.restype = C.c_int
lib.tcl_vidtome_match.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.tcl_vidtome_match.restype = C.c_int
lib.tcl_vidtome_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
lib.tcl_vidtome_plan.restype = C.c_int
lib.tcl_gather_rows.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int,
                                C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tcl_gather_rows.restype = C.c_int

lib.tcl_gemv.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                         C.c_void_p]
lib.tcl_gemv.restype = C.c_int

lib.tcl_adain_blend.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
lib.tcl_adain_blend.restype = C.c_int
lib.tcl_scale_inplace.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]
lib.tcl_scale_inplace.restype = C.c_int
lib.tcl_dpm_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                             C.c_void_p]
lib.tcl_dpm_step.restype = C.c_int


TCL_POSTOPT_MAX_BATCH = 32


class PostoptCtx(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("edited", C.c_void_p), ("past_flows", C.c_void_p), ("mask_bwd", C.c_void_p), ("ypyr", C.c_void_p),
        ("lambda_dssim", C.c_float), ("lambda_flow", C.c_float), ("lambda_tv", C.c_float),
        ("max_batch", C.c_int32), ("norm_batch", C.c_int32), ("norm_valid", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


lib.tcl_postopt_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
lib.tcl_postopt_workspace_bytes.restype = C.c_size_t
lib.tcl_postopt_pyramid_elems.argtypes = [C.c_int, C.c_int]
lib.tcl_postopt_pyramid_elems.restype = C.c_longlong
lib.tcl_postopt_target_elems.argtypes = [C.c_int, C.c_int]
lib.tcl_postopt_target_elems.restype = C.c_longlong
lib.tcl_postopt_build_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_postopt_build_pyramid.restype = C.c_int
lib.tcl_uvt_iteration.argtypes = [C.POINTER(PostoptCtx), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_longlong,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_uvt_iteration.restype = C.c_int
lib.tcl_exposure_iteration.argtypes = [C.POINTER(PostoptCtx), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                       C.c_void_p, C.c_void_p]
lib.tcl_exposure_iteration.restype = C.c_int
lib.tcl_uvt_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                             C.c_void_p]
lib.tcl_uvt_init.restype = C.c_int
lib.tcl_uvt_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_uvt_render.restype = C.c_int
lib.tcl_exposure_bake.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.tcl_exposure_bake.restype = C.c_int

TCL_MAX_RANKS = 8


class UvtShards(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("rows_per_rank", C.c_int64),
        ("fdc", C.c_void_p * TCL_MAX_RANKS), ("grad", C.c_void_p * TCL_MAX_RANKS),
    ]


lib.tcl_uvt_gradient_sharded.argtypes = [C.POINTER(PostoptCtx), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.POINTER(UvtShards),
                                         C.c_void_p, C.c_void_p]
lib.tcl_uvt_gradient_sharded.restype = C.c_int
lib.tcl_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
lib.tcl_peer_alloc.restype = C.c_int
lib.tcl_peer_free.argtypes = [C.c_void_p]
lib.tcl_peer_free.restype = C.c_int
lib.tcl_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
lib.tcl_ipc_open.restype = C.c_int
lib.tcl_ipc_close.argtypes = [C.c_void_p]
lib.tcl_ipc_close.restype = C.c_int
lib.tcl_ipc_close_all.argtypes = []
lib.tcl_ipc_close_all.restype = C.c_int
lib.tcl_peer_barrier.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.tcl_peer_barrier.restype = C.c_int
lib.tcl_peer_barrier_timeouts.argtypes = []
lib.tcl_peer_barrier_timeouts.restype = C.c_longlong
lib.tcl_exposure_gradient.argtypes = [C.POINTER(PostoptCtx), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
lib.tcl_exposure_gradient.restype = C.c_int
lib.tcl_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float,
                              C.c_float, C.c_int, C.c_void_p]
lib.tcl_adam_step.restype = C.c_int



def load_tuning_lib():
    """libtclight_tuning.so: the attention kernel built with its tuning variants and debug hooks (include/tclight_tuning.h).
    Test / tool infrastructure only; returns None when it has not been built."""
    path = os.path.join(_HERE, "libtclight_tuning.so")
    if not os.path.exists(path):
        return None
    t = C.CDLL(path)
    t.tcl_attention.argtypes, t.tcl_attention.restype = lib.tcl_attention.argtypes, C.c_int
    t.tcl_last_error.restype = C.c_char_p
    t.tcl_debug_attention_variant.argtypes, t.tcl_debug_attention_variant.restype = [C.c_int], C.c_int
    t.tcl_debug_attention_trim.argtypes, t.tcl_debug_attention_trim.restype = [C.c_int], C.c_int
    t.tcl_debug_attention_trace.argtypes, t.tcl_debug_attention_trace.restype = [C.c_void_p], None
    return t

lib.tcl_ddim_next.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float,
                              C.c_float, C.c_void_p]
lib.tcl_ddim_next.restype = C.c_int

lib.tcl_warp_bicubic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_warp_bicubic.restype = C.c_int
lib.tcl_max_f32.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
lib.tcl_max_f32.restype = C.c_int
lib.tcl_soft_mask_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                  C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tcl_soft_mask_bwd.restype = C.c_int
lib.tcl_flow_ids_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
lib.tcl_flow_ids_workspace_bytes.restype = C.c_size_t
lib.tcl_flow_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.tcl_flow_ids.restype = C.c_int
lib.tcl_unique_inverse_workspace_bytes.argtypes = [C.c_longlong]
lib.tcl_unique_inverse_workspace_bytes.restype = C.c_size_t
lib.tcl_unique_inverse.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]
lib.tcl_unique_inverse.restype = C.c_int
lib.tcl_adam_step_uvt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_int, C.c_void_p]
lib.tcl_adam_step_uvt.restype = C.c_int

lib.tcl_softmax_rows.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
lib.tcl_softmax_rows.restype = C.c_int
lib.tcl_image_to_nhwc.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p]
lib.tcl_image_to_nhwc.restype = C.c_int
lib.tcl_nhwc_to_image.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                  C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
lib.tcl_nhwc_to_image.restype = C.c_int
