"""Producer side of the stage-2 optimiser on the GPU: mirrors of the reference's
``utils/flow_utils.py`` (warp_flow :5-16, get_soft_mask_bwds :40-54, get_flowid :56-92) and
``utils/general_utils.voxelization`` (:223-256, the ``voxel_size=None`` branch the video data parser
uses, video_dataparser.py:59), bound to the kernels of csrc/flowid.cu.

Same names, argument meaning and return shapes/dtypes as the reference.  Differences, all in the
direction of determinism: no host synchronisation (the ``.max().item()`` thresholds are reduced on the
device), and where several pixels flow onto the same target pixel the largest source index wins (what
the reference's CPU path does; its CUDA ``index_put_`` picks an unspecified one).
"""
from __future__ import annotations


import torch

from ._lib import TclError, check, lib, require_cuda, stream_ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(dtype=torch.float32).contiguous()


def _device_max(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    check(lib.tcl_max_f32(x.data_ptr(), x.numel(), out.data_ptr(), stream_ptr()), "tcl_max_f32")
    return out


def warp_flow(frames: torch.Tensor, past_flows: torch.Tensor) -> torch.Tensor:
    """flow_utils.py:5-16: bicubic warp of ``frames`` [N,C,H,W] by ``past_flows`` [N,>=2,H,W]."""
    require_cuda(frames, past_flows)
    N, Cc, H, W = frames.shape
    fr = _f32c(frames)
    fl = _f32c(past_flows[:, :2])
    out = torch.empty_like(fr)
    check(lib.tcl_warp_bicubic(fr.data_ptr(), fl.data_ptr(), N, Cc, H, W, out.data_ptr(), stream_ptr()), "tcl_warp_bicubic")
    return out


def get_soft_mask_bwds(org_images, flows, past_flows, alpha=0.1, beta=1e2, diff_threshold=0.1, batch_size=64):
    """flow_utils.py:40-54 -> [N,1,H,W] fp32.  ``batch_size`` only chunked the reference's temporaries."""
    require_cuda(org_images, flows, past_flows)
    N, _, H, W = org_images.shape
    if flows.shape[0] != N or past_flows.shape[0] != N:
        raise TclError("get_soft_mask_bwds: images / flows / past_flows must have the same number of frames")
    img, fw, bw = _f32c(org_images), _f32c(flows[:, :2]), _f32c(past_flows[:, :2])
    out = torch.empty((N, 1, H, W), device=img.device, dtype=torch.float32)
    mx = _device_max(img)
    check(lib.tcl_soft_mask_bwd(img.data_ptr(), fw.data_ptr(), bw.data_ptr(), N, H, W, float(alpha), float(beta),
                                float(diff_threshold), mx.data_ptr(), out.data_ptr(), stream_ptr()), "tcl_soft_mask_bwd")
    return out


def get_flowid(frames, flows, mask_bwds, rgb_threshold=0.01, return_count=False):
    """flow_utils.py:56-92 -> flow_ids [N,H,W] int32 (``return_count``: also the number of ids as a
    0-dim device int64 tensor, so that callers need no ``.max()`` pass)."""
    require_cuda(flows, mask_bwds)
    N, _, H, W = frames.shape
    if N * H * W >= 2 ** 31:
        raise TclError("get_flowid: N*H*W >= 2^31 needs int64 ids (not implemented)")
    fr = _f32c(frames.to(flows.device))
    fw = _f32c(flows[:, :2])
    mk = _f32c(mask_bwds)
    ids = torch.empty((N, H, W), device=fr.device, dtype=torch.int32)
    cnt = torch.empty((), device=fr.device, dtype=torch.int64)
    nbytes = lib.tcl_flow_ids_workspace_bytes(N, H, W)
    ws = torch.empty(nbytes, device=fr.device, dtype=torch.uint8)
    mx = _device_max(fr)
    check(lib.tcl_flow_ids(fr.data_ptr(), fw.data_ptr(), mk.data_ptr(), N, H, W, float(rgb_threshold), mx.data_ptr(),
                           ids.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nbytes, stream_ptr()), "tcl_flow_ids")
    return (ids, cnt) if return_count else ids


def voxelization(flow_ids, in_feats_rgb=None, in_feats_coord=None, voxel_size=None, rgb_vox_size=2 / 255, instance_ids=None,
                 xyz_min=None, contract=False, id_range=None):
    """general_utils.py:223-256 for the configuration the video pipeline uses (``voxel_size=None``, no
    instance ids): the inverse map of ``torch.unique(flow_ids, dim=0, return_inverse=True)`` as int64 [n].
    ``id_range`` (exclusive upper bound of the ids) avoids a device->host ``.max()`` when known."""
    if voxel_size is not None or instance_ids is not None:
        raise TclError("voxelization: only the time-only scatter (voxel_size=None, no instance ids) is implemented")
    require_cuda(flow_ids)
    ids = flow_ids.reshape(-1)
    if ids.dtype != torch.int32:
        if ids.numel() and int(ids.max().item()) >= 2 ** 31:
            raise TclError("voxelization: ids >= 2^31 are not implemented")
        ids = ids.to(torch.int32)
    ids = ids.contiguous()
    n = ids.numel()
    if id_range is None:
        id_range = int(ids.max().item()) + 1
    inv = torch.empty(n, device=ids.device, dtype=torch.int64)
    cnt = torch.empty((), device=ids.device, dtype=torch.int64)
    nbytes = lib.tcl_unique_inverse_workspace_bytes(id_range)
    ws = torch.empty(nbytes, device=ids.device, dtype=torch.uint8)
    check(lib.tcl_unique_inverse(ids.data_ptr(), n, id_range, inv.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nbytes,
                                 stream_ptr()), "tcl_unique_inverse")
    return inv


def build_unq_inv(rgbs, future_flows, past_flows, alpha=0.5, diff_threshold=0.1, rgb_threshold=0.01, flow_model="memflow"):
    """The device-resident tail of ``VideoDataParser.load_data`` / ``load_flow`` (video_dataparser.py:44-62,
    107-108) once the flows exist: soft masks -> flow ids -> unique inverse.
    Returns (mask_bwds [N,1,H,W], unq_inv int64 [N*H*W])."""
    gts = rgbs * 2.0 - 1.0 if flow_model.lower() == "memflow" else rgbs      # video_dataparser.py:78-79
    masks = get_soft_mask_bwds(gts, future_flows, past_flows, alpha=alpha, diff_threshold=diff_threshold)
    ids = get_flowid(rgbs, future_flows, masks, rgb_threshold=rgb_threshold)
    N, H, W = ids.shape
    return masks, voxelization(ids.view(-1, 1), id_range=N * H * W)
