"""B200 mirror of the reference's utils/VidToMe/vidtome/merge.py (live functions only:
``bipartite_soft_matching_randframe`` :20-159 and ``bipartite_soft_matching_2s`` :343-463).

Same names, argument meaning and return convention ``(merge, unmerge, info)``; the arithmetic runs
in libtclight.so: normalise+split, tcgen05 score GEMM with fused row max/argmax (no score
matrix), and row-gather merge/unmerge.  Index selection contract (SURVEY.md §7 hard part 1):
``node_max`` is the fp32 accumulator rounded once to the activation dtype, ``node_idx`` the lowest
index among maxima, and the ranking is ``torch.argsort(node_max, descending=True)`` exactly as the
reference calls it, so equal ``node_max`` gives bit-identical ``unm/src/dst`` indices.

Supported subset (everything TC-Light's configs use): ``merge_mode="replace"``, ``unm_pre == 0``
(chunks of at most ``target_stride`` frames => a single local round).  Anything else raises.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from .. import ops
from .._lib import TclError


def do_nothing(x: torch.Tensor, mode: str = None, **kw):
    return x


class _Matching:
    """Result of one bipartite matching: gather maps + the closures the reference API returns."""

    def __init__(self, n_tokens: int, n_src: int, n_dst: int, r: int, d0: int, merge_map, unmerge_map,
                 node_max=None, node_idx=None, edge=None):
        self.N, self.n_src, self.n_dst, self.r, self.d0 = n_tokens, n_src, n_dst, r, d0
        self.merge_map, self.unmerge_map = merge_map, unmerge_map
        self.node_max, self.node_idx, self.edge = node_max, node_idx, edge
        self.unm_num = n_src - r

    # reference-compatible index views (merge.py:100-108), for tests
    def indices(self):
        unm_idx = self.edge[..., self.r:]
        src_idx = self.edge[..., :self.r]
        dst_idx = torch.gather(self.node_idx, -1, src_idx) % self.n_dst
        return unm_idx, src_idx, dst_idx


def _match(a: torch.Tensor, b: torch.Tensor, ratio: float, align_batch: bool, n_tokens: int, d0: int) -> _Matching:
    B, n_src, _ = a.shape
    n_dst = b.shape[1]
    r = min(n_src, int(n_src * ratio))                      # merge.py:90
    node_max, node_idx = ops.vidtome_match(a, b, align_batch)
    # merge.py:98 — same call as the reference on the same values (16-bit scores held in fp32)
    edge = node_max.to(a.dtype).argsort(dim=-1, descending=True)
    if align_batch:
        mm, um = ops.vidtome_plan(edge, node_idx, n_src, n_dst, r, d0)
    else:
        maps = [ops.vidtome_plan(edge[i].contiguous(), node_idx[i].contiguous(), n_src, n_dst, r, d0) for i in range(B)]
        mm = torch.stack([m[0] for m in maps]).contiguous()
        um = torch.stack([m[1] for m in maps]).contiguous()
    return _Matching(n_tokens, n_src, n_dst, r, d0, mm, um, node_max, node_idx, edge)


def _closures(mt: _Matching, two_src_len=None, unmerge_chunk: int = 0):
    def merge(x: torch.Tensor, mode=None, **kw) -> torch.Tensor:
        if mode not in (None, "replace"):
            raise TclError("only merge_mode='replace' is implemented (TC-Light never uses another)")
        return ops.gather_rows(x.contiguous(), None, mt.merge_map)

    def unmerge(x: torch.Tensor, **kw) -> torch.Tensor:
        um = mt.unmerge_map
        if two_src_len is not None:   # merge.py:459 — return only the requested partition
            um = um[..., :two_src_len] if unmerge_chunk == 0 else um[..., two_src_len:]
            um = um.contiguous()
        return ops.gather_rows(x.contiguous(), None, um)

    merge.matching = mt
    unmerge.matching = mt
    return merge, unmerge


def bipartite_soft_matching_randframe(metric: torch.Tensor, F: int, ratio: float, unm_pre: int,
                                      generator: torch.Generator, target_stride: int = 4,
                                      align_batch: bool = False, merge_mode: str = "replace",
                                      randf: int = None) -> Tuple[Callable, Callable, dict]:
    """reference merge.py:20-159.  ``randf`` may be supplied by a caller that already drew it
    from ``generator`` (compute_merge batches its host reads); otherwise it is drawn here with the
    reference's call (merge.py:57)."""
    B, N, _ = metric.shape
    tnum = (N - unm_pre) // F
    if ratio <= 0:
        return do_nothing, do_nothing, {"unm_num": tnum}
    if merge_mode != "replace":
        raise TclError("only merge_mode='replace' is implemented")
    if unm_pre != 0 or F > target_stride:
        raise TclError("multi-round local merging (chunk_size > target_stride) is not implemented; "
                       "TC-Light uses chunk_size <= 4")
    target_stride = min(target_stride, F)
    if randf is None:
        randf = int(torch.randint(0, target_stride, torch.Size([1]), generator=generator, device=generator.device).item())
    d0, d1 = randf * tnum, (randf + 1) * tnum               # dst = tokens of frame `randf` (merge.py:59-64)
    a, b = ops.normalize_split(metric.contiguous(), None, d0, d1)
    mt = _match(a, b, ratio, align_batch, N, d0)
    merge, unmerge = _closures(mt)
    return merge, unmerge, {"unm_num": mt.unm_num}


def bipartite_soft_matching_2s(metric: torch.Tensor, src_len: int, ratio: float, align_batch: bool,
                               merge_mode: str = "replace", unmerge_chunk: int = 0, parts=None):
    """reference merge.py:343-463.  ``parts=(x0, x1)`` lets compute_merge pass the two halves of
    ``metric`` without materialising the concatenation (patch.py:64-70)."""
    if ratio <= 0:
        return do_nothing, do_nothing
    if merge_mode != "replace":
        raise TclError("only merge_mode='replace' is implemented")
    if parts is None:
        x0, x1 = metric.contiguous(), None
        N = metric.shape[1]
    else:
        x0, x1 = parts
        N = x0.shape[1] + x1.shape[1]
    a, b = ops.normalize_split(x0, x1, src_len, N)           # src = first src_len tokens, dst = rest
    mt = _match(a, b, ratio, align_batch, N, src_len)

    def merge(x: torch.Tensor = None, mode=None, parts=None, **kw):
        if parts is not None:
            return ops.gather_rows(parts[0], parts[1], mt.merge_map)
        return ops.gather_rows(x.contiguous(), None, mt.merge_map)

    _, unmerge = _closures(mt, two_src_len=src_len, unmerge_chunk=unmerge_chunk)
    merge.matching = mt
    return merge, unmerge, {"unm_num": mt.unm_num}
