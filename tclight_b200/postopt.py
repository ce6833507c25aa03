"""B200 mirror of the two-stage temporal-consistency optimiser:

  * ``OptDataset``                      reference utils/dataloader.py:9-42
  * ``exposure_align(gen)``             reference generate.py:354-451 (stage 1)
  * ``unique_tensor_optimization(gen)`` reference generate.py:453-533 (stage 2)
  * ``get_expon_lr_func``               reference utils/general_utils.py:31-64

Same hyper-parameters, iteration order, LR indexing and DataLoader RNG consumption as the
reference (batches are drawn by a real ``torch.utils.data.DataLoader(shuffle=True)`` over the
frame indices, so the global CPU RNG advances exactly as in the reference); each iteration is ONE
call into libtclight.so (tcl_exposure_iteration / tcl_uvt_iteration).  Losses are accumulated on
the device and read back once at the end (the reference's per-iteration ``loss.item()`` sync is
the only behavioural difference).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _lib as L
from ._lib import TclError, check, lib, stream_ptr


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear LR interpolation with optional delay (reference utils/general_utils.py:31-64)."""

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)

    return helper


class OptDataset(torch.utils.data.Dataset):
    """GPU-resident frames / backward flows / soft masks (reference utils/dataloader.py:9-42)."""

    def __init__(self, edited_images, past_flows, mask_bwd, device, dtype=torch.float32):
        super().__init__()
        if dtype != torch.float32:
            raise TclError("the optimiser kernels are fp32 (as the reference's OptDataset default)")
        self.edited_images = edited_images.to(dtype=dtype, device=device).contiguous()
        self.past_flows = past_flows.to(dtype=dtype, device=device).contiguous() if past_flows is not None else None
        self.mask_bwd = mask_bwd.to(device=device, dtype=dtype).contiguous() if mask_bwd is not None else None
        self.device, self.dtype = device, dtype
        if self.edited_images.max() > 1:
            self.edited_images = self.edited_images / 255.0

    def __len__(self):
        return len(self.edited_images)

    def __getitem__(self, idx):
        pre = self.edited_images[idx - 1] if idx > 0 else self.edited_images[idx]
        return idx, self.edited_images[idx], pre, self.past_flows[idx], self.mask_bwd[idx]

    @torch.no_grad()
    def exposure_align(self, exposure):
        N, _, H, W = self.edited_images.shape
        check(lib.tcl_exposure_bake(self.edited_images.data_ptr(), exposure.detach().contiguous().data_ptr(), N, H, W, stream_ptr()),
              "tcl_exposure_bake")


def batch_iterator(n_frames: int, batch_size: int):
    """Index batches exactly as ``DataLoader(dataset, batch_size, shuffle=True)`` draws them
    (generate.py:363-367, 466-470): a fresh iterator per epoch, RandomSampler seeded from the
    global CPU RNG."""
    loader = torch.utils.data.DataLoader(range(n_frames), batch_size=batch_size, shuffle=True)
    return loader


class _Context:
    def __init__(self, ds: OptDataset, lambda_dssim, lambda_flow, lambda_tv, batch):
        if ds.past_flows is None or ds.mask_bwd is None:
            raise TclError("optimiser needs past_flows and mask_bwd")
        e = ds.edited_images
        if not e.is_cuda:
            raise TclError("optimiser needs CUDA tensors (no CPU path)")
        N, _, H, W = e.shape
        if batch > L.TCL_POSTOPT_MAX_BATCH:
            raise TclError(f"post_opt.batch_size {batch} > {L.TCL_POSTOPT_MAX_BATCH}")
        self.N, self.H, self.W = N, H, W
        per = lib.tcl_postopt_pyramid_elems(H, W)
        self.ypyr = torch.empty((N, 3, per), device=e.device, dtype=torch.float32)
        check(lib.tcl_postopt_build_pyramid(e.data_ptr(), N, H, W, self.ypyr.data_ptr(), stream_ptr()), "tcl_postopt_build_pyramid")
        wsb = lib.tcl_postopt_workspace_bytes(H, W, batch)
        self.ws = torch.zeros(wsb, device=e.device, dtype=torch.uint8)   # G_pre must start at zero
        c = L.PostoptCtx()
        c.N, c.H, c.W = N, H, W
        c.edited, c.past_flows, c.mask_bwd = e.data_ptr(), ds.past_flows.data_ptr(), ds.mask_bwd.data_ptr()
        c.ypyr = self.ypyr.data_ptr()
        c.lambda_dssim, c.lambda_flow, c.lambda_tv = lambda_dssim, lambda_flow, lambda_tv
        c.max_batch = batch
        c.workspace, c.workspace_bytes = self.ws.data_ptr(), wsb
        self.c = c


def _idx_array(idxs) -> Tuple[C.Array, int]:
    lst = [int(i) for i in idxs]
    return (C.c_int * len(lst))(*lst), len(lst)


def exposure_align(gen) -> Tuple[torch.Tensor, List[float]]:
    """Stage 1 (generate.py:354-451): per-frame 3x4 affine exposure, Adam(default eps 1e-8)."""
    ds = gen.dataset
    N, _, H, W = ds.edited_images.shape
    Bo = gen.opt_batch_size
    dev = ds.edited_images.device
    total_iters = gen.epochs_exposure * N // Bo
    ctx = _Context(ds, gen.lambda_dssim, gen.lambda_flow, gen.lambda_tv, Bo)
    exposure = torch.eye(3, 4, device=dev)[None].repeat(N, 1, 1).contiguous()
    grad, m, v = (torch.zeros_like(exposure) for _ in range(3))
    lr_fn = get_expon_lr_func(gen.exposure_lr_init, gen.exposure_lr_final, lr_delay_steps=gen.exposure_lr_delay_steps,
                              lr_delay_mult=gen.exposure_lr_delay_mult, max_steps=total_iters)
    n_it = gen.epochs_exposure * ((N + Bo - 1) // Bo)
    losses = torch.zeros((max(n_it, 1), 3), device=dev, dtype=torch.float32)
    step = 0
    loader = batch_iterator(N, Bo)
    for epoch in range(gen.epochs_exposure):
        for i, idxs in enumerate(loader):
            iter_idx = epoch * N // Bo + i + 1                     # generate.py:394
            lr = float(lr_fn(iter_idx))
            arr, nb = _idx_array(idxs)
            step += 1
            check(lib.tcl_exposure_iteration(C.byref(ctx.c), arr, nb, exposure.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                             v.data_ptr(), lr, 0.9, 0.999, 1e-8, step, losses[step - 1].data_ptr(), stream_ptr()),
                  "tcl_exposure_iteration")
    gen._exposure = exposure
    ds.exposure_align(exposure)                                     # generate.py:449
    return ds.edited_images, losses[:step, 0].tolist()


def unique_tensor_optimization(gen) -> Tuple[torch.Tensor, List[float]]:
    """Stage 2 (generate.py:453-533): Unique-Video-Tensor optimisation, dense Adam(eps 1e-15)."""
    ds = gen.dataset
    if gen.epochs <= 0:
        return ds.edited_images, []
    N, _, H, W = ds.edited_images.shape
    Bo = gen.opt_batch_size
    dev = ds.edited_images.device
    unq = gen.data_parser.unq_inv
    if unq is None:
        raise TclError("data_parser.unq_inv is not set (run data_parser.load_data first)")
    if N * H * W >= 2 ** 31:
        raise TclError("N*H*W >= 2^31: int64 ids are not implemented yet")
    ids = unq.to(device=dev, dtype=torch.int32).contiguous()
    U = int(unq.max().item()) + 1
    feature_lr = gen.feature_lr * Bo / N                           # generate.py:474
    ctx = _Context(ds, gen.lambda_dssim, gen.lambda_flow, gen.lambda_tv, Bo)
    fdc = torch.empty((U, 3), device=dev, dtype=torch.float32)
    cnt = torch.empty(U, device=dev, dtype=torch.float32)
    check(lib.tcl_uvt_init(ds.edited_images.data_ptr(), ids.data_ptr(), N, H, W, U, fdc.data_ptr(), cnt.data_ptr(), stream_ptr()),
          "tcl_uvt_init")
    del cnt
    grad, m, v = (torch.zeros_like(fdc) for _ in range(3))
    n_it = gen.epochs * ((N + Bo - 1) // Bo)
    losses = torch.zeros((n_it, 3), device=dev, dtype=torch.float32)
    step = 0
    loader = batch_iterator(N, Bo)
    for epoch in range(gen.epochs):
        for idxs in loader:
            arr, nb = _idx_array(idxs)
            step += 1
            check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, nb, ids.data_ptr(), U, fdc.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                        v.data_ptr(), feature_lr, 0.9, 0.999, 1e-15, step, losses[step - 1].data_ptr(), stream_ptr()),
                  "tcl_uvt_iteration")
    images = torch.empty((N, 3, H, W), device=dev, dtype=torch.float32)
    check(lib.tcl_uvt_render(fdc.data_ptr(), ids.data_ptr(), N, H, W, images.data_ptr(), stream_ptr()), "tcl_uvt_render")
    gen._features_dc = fdc
    return images, losses[:step, 0].tolist()
