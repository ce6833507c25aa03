"""TEST INFRASTRUCTURE — CPU/torch restatement of VidToMe's bipartite soft matching as TC-Light
runs it, written as explicit index algebra (no closures) so it can be compared field by field
with the CUDA path.  Follows reference utils/VidToMe/vidtome/merge.py:41-108 (partition, cosine
scores, per-src best dst, descending rank, top-r merged, `% num_dst` under align_batch),
:119-155 (merge = [unmerged src | dst]; unmerge = dst / unmerged / merged-src<-dst), :343-463
(first `src_len` tokens are src) and patch.py:14-91 (local rounds, global pool, role draw).

Pinned against the reference itself: tests/test_oracle_vs_reference.py runs both on the same
inputs in the build container, and tests/golden/vidtome_*.pt hold reference outputs.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch


def partition_randframe(n_tokens: int, F: int, randf: int, unm_pre: int = 0, target_stride: int = 4):
    """src/dst token rows for the local matcher (merge.py:41-71)."""
    tnum = (n_tokens - unm_pre) // F
    stride = min(target_stride, F)
    pos = torch.arange(n_tokens - unm_pre)
    is_dst = (pos // tnum) % stride == randf
    src_rows = pos[~is_dst] + unm_pre
    dst_rows = torch.cat([pos[is_dst] + unm_pre, torch.arange(unm_pre)])
    return src_rows, dst_rows, tnum


def partition_2s(n_tokens: int, src_len: int):
    pos = torch.arange(n_tokens)
    return pos[:src_len], pos[src_len:]


def match(metric: torch.Tensor, src_rows, dst_rows, ratio: float, align_batch: bool,
          best=None) -> Dict[str, torch.Tensor]:
    """Returns node_max, node_idx (before the modulo), rank order, and the three index sets.
    ``best=(node_max, node_idx)`` substitutes externally computed row maxima (tests use this to
    check the index algebra independently of GEMM rounding)."""
    src_rows = src_rows.to(metric.device)
    dst_rows = dst_rows.to(metric.device)
    n_src, n_dst = src_rows.numel(), dst_rows.numel()
    r = min(n_src, int(n_src * ratio))
    if best is None:
        unit = metric / metric.norm(dim=-1, keepdim=True)
        a, b = unit[:, src_rows], unit[:, dst_rows]
        scores = a @ b.transpose(-1, -2)
        if align_batch:
            scores = torch.cat(list(scores), dim=-1)          # [n_src, B*n_dst]
        node_max, node_idx = scores.max(dim=-1)
    else:
        node_max, node_idx = best
    order = node_max.argsort(dim=-1, descending=True)
    src_idx, unm_idx = order[..., :r], order[..., r:]
    dst_idx = torch.gather(node_idx, -1, src_idx)
    if align_batch:
        dst_idx = dst_idx % n_dst
    return dict(node_max=node_max, node_idx=node_idx, order=order, unm_idx=unm_idx, src_idx=src_idx,
                dst_idx=dst_idx, r=r, n_src=n_src, n_dst=n_dst, src_rows=src_rows, dst_rows=dst_rows)


def _per_batch(idx: torch.Tensor, B: int) -> torch.Tensor:
    return idx.unsqueeze(0).expand(B, -1) if idx.dim() == 1 else idx


def merge_tokens(x: torch.Tensor, mt: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[unmerged src | dst]  (merge.py:119-133, replace mode)."""
    B, _, C = x.shape
    src, dst = x[:, mt["src_rows"]], x[:, mt["dst_rows"]]
    unm = torch.gather(src, 1, _per_batch(mt["unm_idx"], B).unsqueeze(-1).expand(-1, -1, C))
    return torch.cat([unm, dst], dim=1)


def unmerge_tokens(y: torch.Tensor, mt: Dict[str, torch.Tensor], n_tokens: int) -> torch.Tensor:
    """merge.py:135-155: dst back in place, unmerged src back in place, merged src <- its dst."""
    B, _, C = y.shape
    n_unm = mt["unm_idx"].shape[-1]
    unm, dst = y[:, :n_unm], y[:, n_unm:]
    out = torch.zeros(B, n_tokens, C, dtype=y.dtype, device=y.device)
    out[:, mt["dst_rows"]] = dst
    bi = torch.arange(B, device=y.device).unsqueeze(-1)
    out[bi, mt["src_rows"][_per_batch(mt["unm_idx"], B)]] = unm
    picked = torch.gather(dst, 1, _per_batch(mt["dst_idx"], B).unsqueeze(-1).expand(-1, -1, C))
    out[bi, mt["src_rows"][_per_batch(mt["src_idx"], B)]] = picked
    return out


class MergeState:
    """Per-block VidToMe state (module.generator / module.global_tokens in the reference)."""

    def __init__(self, generator: torch.Generator):
        self.generator = generator
        self.global_tokens: Optional[torch.Tensor] = None


def compute_merge(state: MergeState, x: torch.Tensor, size: Tuple[int, int], args: Dict,
                  best_fn=None) -> Tuple[torch.Tensor, callable, Dict]:
    """patch.py:14-91.  Returns (merged_tokens, unmerge_fn, trace) where trace records every
    matching (for index-level comparison)."""
    h, w = size
    downsample = int(math.ceil(math.sqrt((h * w) // x.shape[1])))
    trace: Dict = dict(local=[], glob=None)
    if downsample > args["max_downsample"]:
        return x, (lambda t: t), trace
    B = args["batch_size"]
    F = x.shape[0] // B
    n = x.shape[1]
    g = state.generator
    tokens = x.reshape(B, F * n, x.shape[2])
    stack: List = []
    unm = 0
    curF = F
    while curF > 1:
        n_tok = tokens.shape[1]
        if args["local_merge_ratio"] <= 0:
            break
        randf = int(torch.randint(0, min(args["target_stride"], curF), (1,), generator=g, device=g.device).item())
        src_rows, dst_rows, _ = partition_randframe(n_tok, curF, randf, unm, args["target_stride"])
        best = None if best_fn is None else best_fn(tokens, src_rows, dst_rows, args["align_batch"])
        mt = match(tokens, src_rows, dst_rows, args["local_merge_ratio"], args["align_batch"], best)
        mt["randf"] = randf
        trace["local"].append(mt)
        stack.append((mt, n_tok, None))
        unm += mt["unm_idx"].shape[-1]
        tokens = merge_tokens(tokens, mt)
        curF = (tokens.shape[1] - unm) // n
    merged = tokens
    if args["merge_global"]:
        if state.global_tokens is not None:
            pool = state.global_tokens.to(tokens)
            draw = float(torch.rand(1, generator=g, device=g.device).item())
            if draw > args["global_rand"]:
                src_len, cat, local_chunk = tokens.shape[1], torch.cat([tokens, pool], 1), 0
            else:
                src_len, cat, local_chunk = pool.shape[1], torch.cat([pool, tokens], 1), 1
            src_rows, dst_rows = partition_2s(cat.shape[1], src_len)
            best = None if best_fn is None else best_fn(cat, src_rows, dst_rows, args["align_batch"])
            mt = match(cat, src_rows, dst_rows, args["global_merge_ratio"], args["align_batch"], best)
            mt.update(draw=draw, local_chunk=local_chunk, src_len=src_len)
            trace["glob"] = mt
            merged = merge_tokens(cat, mt)
            stack.append((mt, cat.shape[1], (src_len, local_chunk)))
            back = unmerge_tokens(merged, mt, cat.shape[1])
            state.global_tokens = (back[:, :src_len] if local_chunk == 0 else back[:, src_len:]).detach().clone()
        else:
            state.global_tokens = tokens.detach().clone()

    def unmerge(y: torch.Tensor) -> torch.Tensor:
        for mt_, n_tok_, sel in reversed(stack):
            y = unmerge_tokens(y, mt_, n_tok_)
            if sel is not None:
                y = y[:, :sel[0]] if sel[1] == 0 else y[:, sel[0]:]
        return y.reshape(B * F, n, y.shape[-1])

    return merged, unmerge, trace
