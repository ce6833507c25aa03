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
        if self.edited_images.data_ptr() == edited_images.data_ptr():
            self.edited_images = self.edited_images.clone()      # exposure_align bakes in place: never into the caller's tensor
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
        per = lib.tcl_postopt_target_elems(H, W)
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


def _shard_info(gen):
    """(rank, world) when the optimiser runs data parallel (Generator.set_shard + an initialised process group)."""
    world = int(getattr(gen, "_world", 1))
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise TclError("set_shard(world > 1) needs an initialised torch.distributed process group")
        return int(getattr(gen, "_rank", 0)), world
    return 0, 1


def shard_batch(idxs, rank: int, world: int):
    """This rank's slice of a batch plus the global normalisers: every rank draws the same batches (same CPU RNG state,
    see _sync_cpu_rng) and the loss means stay global.  The batch is SORTED by frame index and cut into contiguous
    slices: UVT row ids are assigned in frame order (flow-tracked pixels inherit, new pixels get fresh consecutive ids),
    so a frame's rows cluster in the row range of shard ~ frame * world / N, and rank r, which owns that range, mostly
    gathers from / reduces into its own memory instead of a peer's."""
    lst = sorted(int(i) for i in idxs)
    n = len(lst)
    return lst[rank * n // world:(rank + 1) * n // world], n, sum(1 for i in lst if i > 0)


def _sync_cpu_rng(device):
    """Data parallel: every rank must draw the same DataLoader permutations, but the global CPU RNG has diverged by
    now (get_chunks(shard_len) consumes a length-dependent amount of it in the sharded denoising passes).  Rank 0's
    state is broadcast and adopted by all ranks, so the DP run consumes RNG exactly as rank 0 / a single-GPU run does."""
    import torch.distributed as dist

    st = torch.get_rng_state().to(device)
    dist.broadcast(st, src=0)
    torch.set_rng_state(st.cpu())


class _PeerBuffer:
    """A zero-filled device allocation of its own (tcl_peer_alloc) that other ranks of the box can map; exposes the
    CUDA array interface so the owner can view it as a torch tensor."""

    def __init__(self, nbytes: int):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        check(lib.tcl_peer_alloc(nbytes, C.byref(ptr), handle), "tcl_peer_alloc")
        self.ptr, self.nbytes, self.handle = ptr.value, nbytes, handle.raw

    def view_f32(self, offset_bytes: int, numel: int, device) -> torch.Tensor:
        holder = type("_Span", (), {})()
        holder.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (self.ptr + offset_bytes, False),
                                           "version": 2}
        holder._keep = self
        return torch.as_tensor(holder, device=device)

    def free(self):
        if self.ptr:
            check(lib.tcl_peer_free(self.ptr), "tcl_peer_free")
            self.ptr = None


class UvtShardTable:
    """The UVT rows [U,3] and their gradient [U,4] sharded by row range over the ranks of one box, every shard mapped
    into every rank (tcl_uvt_shards in include/tclight.h).  Layout of one rank's buffer:
    fdc [rows,3] | grad [rows,4] | barrier slots int32[TCL_MAX_RANKS]."""

    def __init__(self, U: int, rank: int, world: int, device):
        import torch.distributed as dist

        self.U, self.rank, self.world, self.device = U, rank, world, device
        rows = (((U + world - 1) // world) + 3) // 4 * 4
        self.rows = max(rows, 256)
        self._grad_off = self.rows * 12
        self._flag_off = self.rows * 28
        torch.cuda.synchronize(device)
        self.buf = _PeerBuffer(self._flag_off + 4 * L.TCL_MAX_RANKS)
        handles = [None] * world
        dist.all_gather_object(handles, self.buf.handle)
        self._mapped = []
        bases = []
        for r in range(world):
            if r == rank:
                bases.append(self.buf.ptr)
                continue
            base = C.c_void_p()
            check(lib.tcl_ipc_open(handles[r], C.byref(base)), "tcl_ipc_open")
            self._mapped.append(base.value)
            bases.append(base.value)
        t = L.UvtShards()
        t.world, t.rank, t.rows_per_rank = world, rank, self.rows
        self._flags = (C.c_void_p * world)()
        for r, bptr in enumerate(bases):
            t.fdc[r], t.grad[r], self._flags[r] = bptr, bptr + self._grad_off, bptr + self._flag_off
        self.c = t
        self.fdc_local = self.buf.view_f32(0, self.rows * 3, device).view(self.rows, 3)
        self.grad_local = self.buf.view_f32(self._grad_off, self.rows * 4, device).view(self.rows, 4)
        self._epoch = 0
        dist.barrier()          # every rank has mapped every shard before anyone touches one

    def barrier(self):
        self._epoch += 1
        check(lib.tcl_peer_barrier(self._flags, self.world, self.rank, self._epoch, stream_ptr()), "tcl_peer_barrier")

    def load_full(self, fdc_full: torch.Tensor):
        """Copy this rank's row range out of a full [U,3] table."""
        lo = self.rank * self.rows
        hi = min(lo + self.rows, self.U)
        if hi > lo:
            self.fdc_local[:hi - lo].copy_(fdc_full[lo:hi])

    def gather_full(self) -> torch.Tensor:
        """All shards -> a full [U,3] table on every rank (once, for the final render)."""
        import torch.distributed as dist

        full = torch.empty((self.world * self.rows, 3), device=self.device, dtype=torch.float32)
        dist.all_gather_into_tensor(full, self.fdc_local.contiguous())
        return full[:self.U].contiguous()

    def close(self):
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        n_to = lib.tcl_peer_barrier_timeouts()
        dist.barrier()
        for bptr in self._mapped:
            check(lib.tcl_ipc_close(bptr), "tcl_ipc_close")
        self._mapped = []
        dist.barrier()          # nobody still maps this rank's buffer
        self.fdc_local = self.grad_local = None
        self.buf.free()
        if n_to:
            raise TclError(f"tcl_peer_barrier timed out {n_to} time(s): the ranks did not run the same iterations")


def _dp_exposure_iteration(ctx, mine, n_global, n_valid_global, params, grad, m, v, lr, step, loss_row):
    """Stage 1: gradient on the local slice with global normalisers -> all-reduce of the [N,12] gradient -> identical
    Adam on every rank."""
    import torch.distributed as dist

    ctx.c.norm_batch, ctx.c.norm_valid = n_global, n_valid_global
    if mine:
        arr, nb = _idx_array(mine)
        check(lib.tcl_exposure_gradient(C.byref(ctx.c), arr, nb, params.data_ptr(), grad.data_ptr(), loss_row.data_ptr(),
                                        stream_ptr()), "tcl_exposure_gradient")
    dist.all_reduce(grad)
    check(lib.tcl_adam_step(params.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), params.numel(), lr, 0.9, 0.999,
                            1e-8, step, stream_ptr()), "tcl_adam_step")


def _dp_uvt_iteration(ctx, tab: UvtShardTable, mine, n_global, n_valid_global, m, v, ids, lr, step, loss_row, marks=None):
    """Stage 2: gather / scatter against the row shards over peer memory, barrier, Adam on the local shard, barrier.
    ``marks`` (measurement only): a list that receives five CUDA events bracketing the four phases."""
    def mark():
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    ctx.c.norm_batch, ctx.c.norm_valid = n_global, n_valid_global
    mark()
    if mine:
        arr, nb = _idx_array(mine)
        check(lib.tcl_uvt_gradient_sharded(C.byref(ctx.c), arr, nb, ids.data_ptr(), C.byref(tab.c), loss_row.data_ptr(),
                                           stream_ptr()), "tcl_uvt_gradient_sharded")
    mark()
    tab.barrier()
    mark()
    check(lib.tcl_adam_step_uvt(tab.fdc_local.data_ptr(), tab.grad_local.data_ptr(), m.data_ptr(), v.data_ptr(), tab.rows, lr,
                                0.9, 0.999, 1e-15, step, stream_ptr()), "tcl_adam_step_uvt")
    mark()
    tab.barrier()
    mark()


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
    rank, world = _shard_info(gen)
    ctx = _Context(ds, gen.lambda_dssim, gen.lambda_flow, gen.lambda_tv, Bo)
    exposure = torch.eye(3, 4, device=dev)[None].repeat(N, 1, 1).contiguous()
    grad, m, v = (torch.zeros_like(exposure) for _ in range(3))
    lr_fn = get_expon_lr_func(gen.exposure_lr_init, gen.exposure_lr_final, lr_delay_steps=gen.exposure_lr_delay_steps,
                              lr_delay_mult=gen.exposure_lr_delay_mult, max_steps=total_iters)
    n_it = gen.epochs_exposure * ((N + Bo - 1) // Bo)
    losses = torch.zeros((max(n_it, 1), 3), device=dev, dtype=torch.float32)
    step = 0
    if world > 1:
        _sync_cpu_rng(dev)
    loader = batch_iterator(N, Bo)
    for epoch in range(gen.epochs_exposure):
        for i, idxs in enumerate(loader):
            iter_idx = epoch * N // Bo + i + 1                     # generate.py:394
            lr = float(lr_fn(iter_idx))
            step += 1
            if world > 1:
                mine, ng, nv = shard_batch(idxs, rank, world)
                _dp_exposure_iteration(ctx, mine, ng, nv, exposure, grad, m, v, lr, step, losses[step - 1])
                continue
            arr, nb = _idx_array(idxs)
            check(lib.tcl_exposure_iteration(C.byref(ctx.c), arr, nb, exposure.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                             v.data_ptr(), lr, 0.9, 0.999, 1e-8, step, losses[step - 1].data_ptr(), stream_ptr()),
                  "tcl_exposure_iteration")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(losses)          # per-rank loss shares are additive
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
    rank, world = _shard_info(gen)
    ctx = _Context(ds, gen.lambda_dssim, gen.lambda_flow, gen.lambda_tv, Bo)
    fdc = torch.empty((U, 3), device=dev, dtype=torch.float32)
    cnt = torch.empty(U, device=dev, dtype=torch.float32)
    check(lib.tcl_uvt_init(ds.edited_images.data_ptr(), ids.data_ptr(), N, H, W, U, fdc.data_ptr(), cnt.data_ptr(), stream_ptr()),
          "tcl_uvt_init")
    del cnt
    tab = None
    if world > 1:
        tab = UvtShardTable(U, rank, world, dev)
        tab.load_full(fdc)
        del fdc
        m, v = torch.zeros((tab.rows, 3), device=dev), torch.zeros((tab.rows, 3), device=dev)
        _sync_cpu_rng(dev)
        tab.barrier()
    else:
        m, v = torch.zeros_like(fdc), torch.zeros_like(fdc)
        grad = torch.zeros((U, 4), device=dev, dtype=torch.float32)     # {dR, dG, dB, pad}: one 16-byte reduction per scatter
    n_it = gen.epochs * ((N + Bo - 1) // Bo)
    losses = torch.zeros((n_it, 3), device=dev, dtype=torch.float32)
    step = 0
    loader = batch_iterator(N, Bo)
    try:
        for epoch in range(gen.epochs):
            for idxs in loader:
                step += 1
                if world > 1:
                    mine, ng, nv = shard_batch(idxs, rank, world)
                    _dp_uvt_iteration(ctx, tab, mine, ng, nv, m, v, ids, feature_lr, step, losses[step - 1])
                    continue
                arr, nb = _idx_array(idxs)
                check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, nb, ids.data_ptr(), U, fdc.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                            v.data_ptr(), feature_lr, 0.9, 0.999, 1e-15, step, losses[step - 1].data_ptr(), stream_ptr()),
                      "tcl_uvt_iteration")
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(losses)
            fdc = tab.gather_full()
    finally:
        if tab is not None:
            tab.close()          # unmaps the peers' shards and frees this rank's (collective: every rank reaches it)
    images = torch.empty((N, 3, H, W), device=dev, dtype=torch.float32)
    check(lib.tcl_uvt_render(fdc.data_ptr(), ids.data_ptr(), N, H, W, images.data_ptr(), stream_ptr()), "tcl_uvt_render")
    gen._features_dc = fdc
    return images, losses[:step, 0].tolist()


# ---------------------------------------------------------------------------------------------
# synthetic workload + measurement helpers (bench.py, __graft_entry__.smoke)
# ---------------------------------------------------------------------------------------------
def synthetic_workload(n_frames: int, H: int, W: int, device, track_len: int = 10, seed: int = 0):
    """Seeded synthetic stage-2 inputs of the named shape, generated on the device with torch ops
    (test data, not the product path): a smooth texture translating by (2, 1) px/frame, noisy
    'edited' frames, backward flows (-2, -1) + 0.25 px wobble, soft masks, and flow-tracked ids whose
    tracks are cut every `track_len` frames (U / (N*H*W) ~ 1/track_len, in the 6-13 % range the
    reference's voxelization produces on real clips)."""
    g = torch.Generator(device=device).manual_seed(seed)
    dx, dy = 2, 1
    Hb, Wb = H + n_frames * dy, W + n_frames * dx
    coarse = torch.rand(1, 3, Hb // 16 + 2, Wb // 16 + 2, device=device, generator=g)
    big = torch.nn.functional.interpolate(coarse, size=(Hb, Wb), mode="bilinear", align_corners=False)[0]
    edited = torch.empty((n_frames, 3, H, W), device=device, dtype=torch.float32)
    ids = torch.empty((n_frames, H, W), device=device, dtype=torch.int32)
    yy = torch.arange(H, device=device, dtype=torch.int64)[:, None]
    xx = torch.arange(W, device=device, dtype=torch.int64)[None, :]
    for f in range(n_frames):
        oy, ox = (n_frames - 1 - f) * dy, (n_frames - 1 - f) * dx
        edited[f] = (big[:, oy:oy + H, ox:ox + W] * 0.8 + 0.1 + 0.02 * torch.randn(3, H, W, device=device, generator=g)).clamp_(0, 1)
        ids[f] = ((f // track_len) * (Hb * Wb) + (yy + oy) * Wb + (xx + ox)).to(torch.int32)
    # compact the id space (what torch.unique(return_inverse) does in the reference's voxelization)
    flat = ids.reshape(-1).long()
    present = torch.zeros(int(flat.max().item()) + 1, device=device, dtype=torch.bool)
    present[flat] = True
    remap = torch.cumsum(present.to(torch.int32), 0, dtype=torch.int32) - 1
    unq_inv = remap[flat].to(torch.int64)
    del present, remap, flat, ids
    flows = torch.empty((n_frames, 2, H, W), device=device, dtype=torch.float32)
    flows[:, 0] = -dx + 0.25 * torch.sin(torch.arange(W, device=device).float() / 17.0)[None, None, :]
    flows[:, 1] = -dy + 0.25 * torch.cos(torch.arange(H, device=device).float() / 13.0)[None, :, None]
    masks = 0.5 + 0.5 * torch.rand((n_frames, 1, H, W), device=device, generator=g)
    masks[:, :, :, :dx + 1] = 0
    masks[:, :, :dy + 1, :] = 0
    return edited, flows, masks, unq_inv


def bench_postopt(device, n_frames: int, H: int, W: int, iters: int = 20, rank: int = 0, world: int = 1, batch: int = 16,
                  traffic=None, full_budget: bool = True):
    """Stage-2 / stage-1 measurements at the named shape for the bench JSON line: (1) `iters` stage-2 iterations (16
    frames each) timed with CUDA events; (2) the reference's full budget (70 epochs of stage 2, 35 of stage 1) through
    the public API, loss curve end points.  Algorithmic bytes per stage-2 iteration: 80*Bo*P + 84*U.
    Returns (stage2 dict, stage1 dict)."""
    import json
    import os
    import time
    import types

    edited, flows, masks, unq_inv = synthetic_workload(n_frames, H, W, device)
    ds = OptDataset(edited, flows, masks, device=device)
    del edited, flows, masks
    N = n_frames
    Bo = batch
    ids = unq_inv.to(torch.int32).contiguous()
    U = int(unq_inv.max().item()) + 1
    ctx = _Context(ds, 0.2, 0.8, 0.05, Bo)
    fdc = torch.empty((U, 3), device=device, dtype=torch.float32)
    cnt = torch.empty(U, device=device, dtype=torch.float32)
    check(lib.tcl_uvt_init(ds.edited_images.data_ptr(), ids.data_ptr(), N, H, W, U, fdc.data_ptr(), cnt.data_ptr(), stream_ptr()), "tcl_uvt_init")
    del cnt
    tab = None
    if world > 1:
        tab = UvtShardTable(U, rank, world, device)
        tab.load_full(fdc)
        del fdc
        m, v = torch.zeros((tab.rows, 3), device=device), torch.zeros((tab.rows, 3), device=device)
        tab.barrier()
    else:
        m, v = torch.zeros_like(fdc), torch.zeros_like(fdc)
        grad = torch.zeros((U, 4), device=device, dtype=torch.float32)
    losses = torch.zeros((iters + 8, 3), device=device)
    gcpu = torch.Generator().manual_seed(0)
    lr = 0.05 * Bo / N

    def run(k, step0):
        for i in range(k):
            idxs = torch.randperm(N, generator=gcpu)[:Bo].tolist()      # same seed => same batches on every rank
            row = losses[(step0 + i) % len(losses)]
            if world > 1:
                mine, ng, nv = shard_batch(idxs, rank, world)
                _dp_uvt_iteration(ctx, tab, mine, ng, nv, m, v, ids, lr, step0 + i + 1, row, marks=phase_marks)
                continue
            arr, nb = _idx_array(idxs)
            check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, nb, ids.data_ptr(), U, fdc.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                        v.data_ptr(), lr, 0.9, 0.999, 1e-15, step0 + i + 1, row.data_ptr(), stream_ptr()),
                  "tcl_uvt_iteration")

    phase_marks = None

    def sync():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    run(3, 0)
    sync()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run(iters, 3)
    e.record()
    sync()
    ms = max_over_ranks(s.elapsed_time(e) / iters)
    phases = None
    if world > 1:
        import torch.distributed as dist
        # a second, instrumented pass: events around the four phases of an iteration (this rank's view, mean over iterations)
        phase_marks = []
        run(min(iters, 10), 3 + iters)
        sync()
        k = len(phase_marks) // 5
        acc = [0.0] * 4
        for j in range(k):
            ev = phase_marks[5 * j:5 * j + 5]
            for q_ in range(4):
                acc[q_] += ev[q_].elapsed_time(ev[q_ + 1])
        mine_ms = torch.tensor([a / max(k, 1) for a in acc], device=device, dtype=torch.float64)
        mx = mine_ms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = mine_ms.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        names = ["gradient_kernels", "barrier_after_scatter", "adam_local_shard", "barrier_after_adam"]
        phases = {n: {"min_ms": mn[i].item(), "max_ms": mx[i].item()} for i, n in enumerate(names)}
        phase_marks = None
        dist.all_reduce(losses)
        tab.close()
    loss_fl = [losses[3, 0].item(), losses[(3 + iters - 1) % len(losses), 0].item()]
    del m, v, tab, ctx
    if world == 1:
        del fdc, grad
    torch.cuda.empty_cache()
    P_ = H * W
    alg_bytes = 80.0 * Bo * P_ + 84.0 * U
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["hbm_gbs"]
        kind = "measured"
    except Exception:
        peak, kind = 6650.0, "fallback"
    ach = alg_bytes / (ms * 1e-3) / 1e9
    traffic = traffic or {}
    s2 = {"metric": "stage2_iters_per_sec", "value": 1e3 / ms, "unit": "it/s", "ms_per_iter": ms, "frames": N, "batch": Bo,
          "U": U, "U_over_NP": U / (N * P_), "algorithmic_bytes_per_iter": alg_bytes,
          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": kind, "unit": "GB/s", "frac": ach / peak,
                       "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
                       "traffic_unit": "bytes/iteration (ncu dram__bytes_read+write summed over the iteration's kernels, 1 GPU)"},
          "loss_first_last": loss_fl, "n_gpus": world, "phase_ms_over_ranks": phases,
          "parallelism": (f"batch-parallel x{world}; UVT rows + gradient sharded by row range, gathered / reduced through peer memory over "
                          "NVLink inside the gather and level-0 kernels, Adam on the local shard") if world > 1 else "single GPU",
          "note": "dense-Adam semantics (every UVT row updated every iteration, as torch.optim.Adam does)"}
    s1 = {}
    if full_budget:
        gen = types.SimpleNamespace(dataset=ds, data_parser=types.SimpleNamespace(unq_inv=unq_inv), lambda_dssim=0.2, lambda_flow=0.8,
                                    lambda_tv=0.05, epochs_exposure=35, epochs=70, opt_batch_size=Bo, feature_lr=0.05,
                                    exposure_lr_init=0.01, exposure_lr_final=0.001, exposure_lr_delay_steps=0,
                                    exposure_lr_delay_mult=0.0, _world=world, _rank=rank)
        torch.manual_seed(12345)
        sync()
        t0 = time.perf_counter()
        _, curve = unique_tensor_optimization(gen)
        sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        s2["full_budget"] = {"epochs": 70, "iterations": len(curve), "seconds": dt, "iters_per_sec": len(curve) / dt,
                             "loss_first": curve[0], "loss_last": curve[-1], "loss_min": min(curve),
                             "note": "Generator.unique_tensor_optimization: the reference's iteration budget (generate.py:453-533) through the "
                                     "public API, DataLoader draws and UVT init / final render included (wall clock)"}
        torch.manual_seed(12345)
        sync()
        t0 = time.perf_counter()
        _, curve1 = exposure_align(gen)
        sync()
        dt1 = max_over_ranks(time.perf_counter() - t0)
        s1 = {"metric": "stage1_iters_per_sec", "value": len(curve1) / dt1, "unit": "it/s", "epochs": 35, "iterations": len(curve1),
              "seconds": dt1, "loss_first": curve1[0], "loss_last": curve1[-1], "n_gpus": world,
              "parallelism": f"batch-parallel x{world}, NCCL all-reduce of the [N,12] exposure gradient" if world > 1 else "single GPU",
              "note": "Generator.exposure_align (generate.py:354-451) through the public API (wall clock)"}
    return s2, s1


def bench_stage2(device, n_frames: int, H: int, W: int, iters: int = 20, rank: int = 0, world: int = 1, batch: int = 16):
    return bench_postopt(device, n_frames, H, W, iters=iters, rank=rank, world=world, batch=batch, full_budget=False)[0]
