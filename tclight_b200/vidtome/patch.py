"""B200 mirror of the reference's utils/VidToMe/vidtome/patch.py: ``compute_merge`` (:14-91) and the
patching API ``apply_patch / remove_patch / update_patch / collect_from_patch`` (:234-387).

Differences from the reference, none of which change results:
  * the global-token pool stays on the device (the reference round-trips it through the CPU,
    patch.py:80, 82);
  * the two per-call random draws (``randint`` for the target frame, merge.py:57; ``rand`` for the
    src/dst role, patch.py:62) are drawn from the module generator with the reference's calls and
    order, but read back to the host with a single synchronisation;
  * local+global unmerge are composed into ONE gather map so the block applies
    ``unmerge(attn_out) + residual`` in a single pass (patch.py:168-169).
The UNet that owns the patched blocks is ``tclight_b200.unet.UNetB200``; its transformer blocks are
modules *named* ``BasicTransformerBlock`` so the reference's own ``apply_patch`` would also find them.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, Tuple

import torch

from .._lib import TclError

from . import merge
from .utils import func_warper, init_generator, isinstance_str, join_frame, join_warper, split_warper
from .. import ops


class MergePlan:
    """What a transformer block needs: merged tokens for self-attention and the composed map
    token(position in [B, F*n]) -> row of the merged sequence."""

    def __init__(self, merged_tokens, total_unmerge_map, fsize):
        self.merged_tokens = merged_tokens
        self.total_unmerge_map = total_unmerge_map
        self.fsize = fsize


def compute_merge_plan(module, x: torch.Tensor, tome_info: Dict[str, Any]):
    """Core of compute_merge; returns (m, u, merged_tokens, plan)."""
    original_h, original_w = tome_info["size"]
    original_tokens = original_h * original_w
    downsample = int(math.ceil(math.sqrt(original_tokens // x.shape[1])))
    args = tome_info["args"]
    generator = module.generator
    fsize = x.shape[0] // args["batch_size"]
    tsize = x.shape[1]

    if downsample > args["max_downsample"]:
        module._tome_active = False
        return merge.do_nothing, merge.do_nothing, x, None
    module._tome_active = True

    if args["generator"] is None:
        args["generator"] = init_generator(x.device)
    elif args["generator"].device != x.device:
        args["generator"] = init_generator(x.device, fallback=args["generator"])

    do_local = fsize > 1 and args["local_merge_ratio"] > 0
    has_pool = args["merge_global"] and getattr(module, "global_tokens", None) is not None
    # ---- random draws, reference order: randint inside randframe (merge.py:57), then rand (patch.py:62)
    queue = getattr(module, "_draw_queue", None)
    if queue:
        # values drawn ahead of time by prefetch_draws (same generator, same call order, ONE host sync per pass)
        randf = grand = None
        if do_local:
            kind, bound, val = queue.popleft()
            if kind != "i" or bound != min(args["target_stride"], fsize):
                raise TclError("VidToMe draw queue out of step with the chunk plan (local draw)")
            randf = int(val)
        if has_pool:
            kind, _, val = queue.popleft()
            if kind != "r":
                raise TclError("VidToMe draw queue out of step with the chunk plan (global draw)")
            grand = val
    else:
        draws = []
        if do_local:
            draws.append(torch.randint(0, min(args["target_stride"], fsize), torch.Size([1]), generator=generator,
                                       device=generator.device).to(torch.float32))
        if has_pool:
            draws.append(torch.rand(1, generator=generator, device=generator.device))
        vals = torch.cat(draws).tolist() if draws else []          # host sync: the roles decide tensor shapes
        randf = int(vals[0]) if do_local else None
        grand = vals[-1] if has_pool else None

    local_tokens = join_frame(x, fsize)
    m_ls = [join_warper(fsize)]
    u_ls = [split_warper(fsize)]
    unm = 0
    curF = fsize
    total_map = None          # token position -> row of current merged sequence
    while curF > 1:
        m, u, ret = merge.bipartite_soft_matching_randframe(
            local_tokens, curF, args["local_merge_ratio"], unm, generator, args["target_stride"],
            args["align_batch"], randf=randf)
        unm += ret["unm_num"]
        m_ls.append(m)
        u_ls.append(u)
        if m is merge.do_nothing:
            break
        local_tokens = m(local_tokens)
        total_map = m.matching.unmerge_map
        curF = (local_tokens.shape[1] - unm) // tsize
    merged_tokens = local_tokens

    if args["merge_global"]:
        if has_pool:
            pool = module.global_tokens
            if pool.device != local_tokens.device or pool.dtype != local_tokens.dtype:
                pool = pool.to(local_tokens)
            if grand > args["global_rand"]:
                src_len = local_tokens.shape[1]
                parts = (local_tokens.contiguous(), pool.contiguous())
                local_chunk = 0
            else:
                src_len = pool.shape[1]
                parts = (pool.contiguous(), local_tokens.contiguous())
                local_chunk = 1
            m, u, _ = merge.bipartite_soft_matching_2s(None, src_len, args["global_merge_ratio"], args["align_batch"],
                                                       unmerge_chunk=local_chunk, parts=parts)
            if m is not merge.do_nothing:
                merged_tokens = m(parts=parts)
                m_ls.append(lambda t, _m=m, _p=parts, **kw: _m(parts=_p))
                u_ls.append(u)
                gmap = m.matching.unmerge_map
                gmap_local = (gmap[..., :src_len] if local_chunk == 0 else gmap[..., src_len:]).contiguous()
                # new pool = unmerged local tokens (patch.py:80), kept on device
                module.global_tokens = ops.gather_rows(merged_tokens, None, gmap_local)
                # compose token -> local row -> merged row
                if total_map is None:
                    total_map = gmap_local
                elif gmap_local.dim() == 1 and total_map.dim() == 1:
                    total_map = gmap_local[total_map.long()].contiguous()
                else:
                    g2 = gmap_local if gmap_local.dim() == 2 else gmap_local.expand(x.shape[0] // fsize, -1)
                    t2 = total_map if total_map.dim() == 2 else total_map.expand(x.shape[0] // fsize, -1)
                    total_map = torch.gather(g2, 1, t2.long()).contiguous()
        else:
            module.global_tokens = local_tokens.detach().clone()

    m = func_warper(m_ls)
    u = func_warper(u_ls[::-1])
    plan = MergePlan(merged_tokens, total_map, fsize)
    return m, u, merged_tokens, plan


def compute_merge(module, x: torch.Tensor, tome_info: Dict[str, Any]) -> Tuple[Callable, Callable, torch.Tensor]:
    """reference patch.py:14-91: returns (merge_fn, unmerge_fn, merged_tokens)."""
    m, u, merged, _ = compute_merge_plan(module, x, tome_info)
    return m, u, merged


# ---------------------------------------------------------------------------------------------
# patching API (reference patch.py:234-387)
# ---------------------------------------------------------------------------------------------
def _unet_of(model):
    return model.unet if hasattr(model, "unet") else model


def apply_patch(model, local_merge_ratio: float = 0.9, merge_global: bool = False, global_merge_ratio=0.8,
                max_downsample: int = 2, seed: int = 123, batch_size: int = 2, include_control: bool = False,
                align_batch: bool = False, target_stride: int = 4, global_rand=0.5):
    """Same signature and defaults as reference patch.py:234-245.  ``model`` is a pipeline-like
    object with ``.unet`` or the UNet itself; the UNet must expose ``named_modules()`` with blocks
    named ``BasicTransformerBlock`` (UNetB200 does)."""
    remove_patch(model)
    unet = _unet_of(model)
    if not hasattr(unet, "named_modules"):
        raise RuntimeError("Provided model was not a Stable Diffusion / Latent Diffusion model, as expected.")
    unet._tome_info = {
        "size": None,
        "hooks": [],
        "args": {
            "max_downsample": max_downsample, "generator": None, "seed": seed, "batch_size": batch_size,
            "align_batch": align_batch, "merge_global": merge_global, "global_merge_ratio": global_merge_ratio,
            "local_merge_ratio": local_merge_ratio, "global_rand": global_rand, "target_stride": target_stride,
        },
    }
    for _, module in unet.named_modules():
        if isinstance_str(module, "BasicTransformerBlock"):
            module._tome_info = unet._tome_info
            module._tome_patched = True
    return model


def prefetch_draws(model, fsizes) -> int:
    """Draws, ahead of time, every random number the patched blocks will consume during a pass whose chunk-forwards
    have ``fsizes`` frames each (reference order per module generator: ``randint`` for the target frame when the
    chunk is merged locally, merge.py:57, then ``rand`` for the global-merge role when a pool exists, patch.py:62),
    and fetches them with ONE device->host copy.  Without this every block of every chunk-forward synchronises the
    host (2 750 times per 300-frame step), which leaves the GPU idle while Python refills the launch queue.  The
    generators see exactly the same calls in the same order, so the values — and therefore the merge indices — are
    unchanged.  Blocks that have not run yet (no generator forked, activity unknown) keep drawing directly.
    Returns the number of values prefetched."""
    import collections

    unet = _unet_of(model)
    info = getattr(unet, "_tome_info", None)
    if info is None:
        return 0
    args = info["args"]
    todo = []
    for _, module in unet.named_modules():
        if not getattr(module, "_tome_patched", False) or not getattr(module, "_tome_active", False):
            continue
        gen = getattr(module, "generator", None)
        if gen is None:
            continue
        if getattr(module, "_draw_queue", None):
            raise TclError("VidToMe draw queue not empty at the start of a pass")
        has_pool = bool(args["merge_global"]) and getattr(module, "global_tokens", None) is not None
        q = []
        for fs in fsizes:
            if fs > 1 and args["local_merge_ratio"] > 0:
                bound = min(args["target_stride"], fs)
                q.append(["i", bound, torch.randint(0, bound, torch.Size([1]), generator=gen, device=gen.device).to(torch.float32)])
            if has_pool:
                q.append(["r", 0, torch.rand(1, generator=gen, device=gen.device)])
            if args["merge_global"]:
                has_pool = True
        module._draw_queue = q
        todo += q
    if todo:
        vals = torch.cat([t[2].reshape(1) for t in todo]).tolist()
        for t, v in zip(todo, vals):
            t[2] = v
    for _, module in unet.named_modules():
        q = getattr(module, "_draw_queue", None)
        if isinstance(q, list):
            module._draw_queue = collections.deque(q)
    return len(todo)


def assert_draws_consumed(model) -> None:
    unet = _unet_of(model)
    for name, module in unet.named_modules():
        if getattr(module, "_draw_queue", None):
            raise TclError(f"VidToMe draw queue of {name} has {len(module._draw_queue)} unconsumed values")


def remove_patch(model):
    unet = _unet_of(model)
    if not hasattr(unet, "named_modules"):
        return model
    for _, module in unet.named_modules():
        if hasattr(module, "_tome_info"):
            del module._tome_info
        if hasattr(module, "_tome_patched"):
            module._tome_patched = False
    return model


def update_patch(model, **kwargs):
    """reference patch.py:358-370: setattr on every patched module (e.g. global_tokens=None)."""
    unet = _unet_of(model)
    for _, module in unet.named_modules():
        if hasattr(module, "_tome_info"):
            for k, v in kwargs.items():
                setattr(module, k, v)
    return model


def collect_from_patch(model, attr="tome"):
    unet = _unet_of(model)
    return {name: getattr(module, attr) for name, module in unet.named_modules() if hasattr(module, attr)}
