"""TEST INFRASTRUCTURE — CPU/torch restatement of TC-Light's multi-axis sampler loop as plain
functions: chunk planning (reference utils/VidToMe/generate_utils.py:174-205), per-chunk CFG noise
prediction (generate.py:287-352), the yt-plane pass with overlapping windows, AdaIN and the
variance-preserving blend (generate.py:241-284, utils/general_utils.py:137-156), scheduler step and
pool reset (generate.py:216-237).  Runs on any device with the oracle UNet (oracle/unet_ref.py +
apply_oracle_patch) and oracle scheduler.  Pinned to the reference's own ``Generator.ddim_sample``
by tests/test_oracle_vs_reference.py (same seeds => identical tensors on CPU).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from .scheduler_ref import DPMSolverSDEKarras
from .unet_ref import apply_oracle_patch, reset_oracle_pool


def plan_chunks(length: int, chunk_size: int = 4, merge_global: bool = True, chunk_ord: str = "mix", perm_div: float = 4.0):
    """Random partition of range(length); consumes np.random.randint, np.random.rand and
    torch.randperm in the reference's order."""
    idx = torch.arange(length)
    first = np.random.randint(0, chunk_size) + 1
    rest = idx[first:].split(chunk_size, dim=0)
    parts = [idx[:first]] + list(rest) if len(rest[0]) > 0 else [idx[:first]]
    if np.random.rand() > 0.5:
        parts = parts[::-1]
    if not merge_global:
        return parts
    if chunk_ord == "rand":
        order = torch.randperm(len(parts)).tolist()
    elif chunk_ord == "mix":
        perm = torch.randperm(len(parts)).tolist()
        k = int(len(perm) / perm_div)
        tail = sorted(perm[k:])
        if k > 0:
            head = perm[:k]
            if abs(tail[-1] - head[-1]) < abs(tail[0] - head[-1]):
                tail = tail[::-1]
            order = head + tail
        else:
            order = tail
    else:
        order = list(range(len(parts)))
    return [parts[i] for i in order]


def window_plan(num_frames: int, win: int):
    n = math.ceil((num_frames - 1) / (win - 1))
    if n <= 1:
        return [0], [0]
    total = n * win - num_frames
    ov = total // (n - 1)
    overlaps = [ov] * (n - 2) + [ov + total % (n - 1)]
    cum = np.cumsum(overlaps)
    return [0] + [int((i + 1) * win - cum[i]) for i in range(n - 1)], overlaps


def cfg_noise(unet, x, text2, t, cond, guidance_scale):
    F = x.shape[0]
    ehs = text2.repeat_interleave(F, dim=0)
    eps = unet(torch.cat([x, x]), t, encoder_hidden_states=ehs, cross_attention_kwargs={"concat_conds": cond}).sample
    un, co = eps.chunk(2)
    return un + guidance_scale * (co - un)


def adain(content, style, eps=1e-5):
    def stats(f):
        n, c = f.shape[:2]
        flat = f.reshape(n, c, -1)
        return flat.mean(dim=2).view(n, c, 1, 1), (flat.var(dim=2) + eps).sqrt().view(n, c, 1, 1)

    sm, ss = stats(style)
    cm, cs = stats(content)
    return (content - cm) / cs * ss + sm


@torch.no_grad()
def ddim_sample_oracle(unet, x, conds, conds_t, concat_conds, n_timesteps=25, alpha_t=0.0, final_factor_t=0.01,
                       win_size_t=64, guidance_scale=2.0, chunk_size=4, merge_global=True, chunk_ord="mix-4",
                       local_merge_ratio=0.6, global_merge_ratio=0.5, global_rand=0.5, rng=None, patch=True):
    if patch and not hasattr(unet, "_oracle_tome"):
        apply_oracle_patch(unet, local_merge_ratio, merge_global, global_merge_ratio, global_rand=global_rand)
    perm_div = float(chunk_ord.split("-")[-1]) if "-" in chunk_ord else 3.0
    ord_kind = "mix" if "mix" in chunk_ord else chunk_ord
    sched = DPMSolverSDEKarras()
    sched.set_timesteps(n_timesteps, device=x.device)
    noises = torch.zeros_like(x)
    noises_t = torch.zeros_like(x)
    steps = sched.timesteps
    for i, t in enumerate(steps):
        for part in plan_chunks(len(x), chunk_size, merge_global, ord_kind, perm_div):
            cc = concat_conds[part] if concat_conds is not None else None
            noises[part] = cfg_noise(unet, x[part], conds, t, cc, guidance_scale)
        if alpha_t > 0:
            a = alpha_t * final_factor_t ** min(i / len(steps), 1)
            starts, overlaps = window_plan(len(x), win_size_t)
            cols = plan_chunks(x.shape[-1], chunk_size, merge_global, ord_kind, perm_div)
            for wi, s in enumerate(starts):
                for part in cols:
                    xt = x[s:s + win_size_t][:, :, :, part].permute(3, 1, 0, 2)
                    ct = concat_conds[s:s + win_size_t][:, :, :, part].permute(3, 1, 0, 2) if concat_conds is not None else None
                    pred = cfg_noise(unet, xt, conds_t, t, ct, guidance_scale)
                    noises_t[s:s + win_size_t, :, :, part] = pred.permute(2, 1, 3, 0)
                if s > 0:
                    noises_t[s:s + overlaps[wi - 1]] *= np.sqrt(0.5)
            noises_t = adain(noises_t, noises)
            noises = (a ** 0.5) * noises_t + ((1 - a) ** 0.5) * noises
        x = sched.step(noises, t, x, generator=rng)[0]
        if merge_global:
            reset_oracle_pool(unet)
    return x


class RankRng:
    """Host RNG streams of one emulated rank (np.random + torch CPU): every process of a multi-GPU run owns its
    own streams, all seeded identically, and consumes them at its own pace."""

    def __init__(self, seed: int):
        self.np_state = np.random.RandomState(seed).get_state()
        g = torch.Generator().manual_seed(seed)
        self.torch_state = g.get_state()

    def __enter__(self):
        self._saved = (np.random.get_state(), torch.get_rng_state())
        np.random.set_state(self.np_state)
        torch.set_rng_state(self.torch_state)
        return self

    def __exit__(self, *a):
        self.np_state, self.torch_state = np.random.get_state(), torch.get_rng_state()
        np.random.set_state(self._saved[0])
        torch.set_rng_state(self._saved[1])
        return False


def shard_range(n: int, rank: int, world: int):
    return (n * rank) // world, (n * (rank + 1)) // world


@torch.no_grad()
def ddim_sample_oracle_sharded(unets, x, conds, conds_t, concat_conds, world, n_timesteps=25, alpha_t=0.0,
                               final_factor_t=0.01, win_size_t=64, guidance_scale=2.0, chunk_size=4, chunk_ord="mix-4",
                               local_merge_ratio=0.6, global_merge_ratio=0.5, global_rand=0.5, rng=None, seed=12345):
    """The multi-GPU semantics of SURVEY.md §8(e): "the reference with the global-token pool reset at shard
    boundaries".  Rank r of `world` runs the reference loop body (generate.py:216-237) on frames
    [N r / world, N (r+1) / world) in the xy pass and on latent columns [W r / world, W (r+1) / world) in the yt pass,
    with its OWN VidToMe pool / module generators (`unets[r]`: one patched oracle UNet per rank, identical weights)
    and its own host RNG streams (seeded alike); AdaIN, blend and the scheduler step are replicated.  world == 1 is
    `ddim_sample_oracle`."""
    assert len(unets) == world
    for u in unets:
        if not hasattr(u, "_oracle_tome"):
            apply_oracle_patch(u, local_merge_ratio, True, global_merge_ratio, global_rand=global_rand)
    perm_div = float(chunk_ord.split("-")[-1]) if "-" in chunk_ord else 3.0
    ord_kind = "mix" if "mix" in chunk_ord else chunk_ord
    rngs = [RankRng(seed) for _ in range(world)]
    sched = DPMSolverSDEKarras()
    sched.set_timesteps(n_timesteps, device=x.device)
    steps = sched.timesteps
    N, W = len(x), x.shape[-1]
    for i, t in enumerate(steps):
        noises = torch.zeros_like(x)
        for r in range(world):
            f0, f1 = shard_range(N, r, world)
            with rngs[r]:
                for part in plan_chunks(f1 - f0, chunk_size, True, ord_kind, perm_div):
                    part = part + f0
                    noises[part] = cfg_noise(unets[r], x[part], conds, t, concat_conds[part], guidance_scale)
        if alpha_t > 0:
            a = alpha_t * final_factor_t ** min(i / len(steps), 1)
            starts, overlaps = window_plan(N, win_size_t)
            noises_t = torch.zeros_like(x)
            for r in range(world):
                w0, w1 = shard_range(W, r, world)
                with rngs[r]:
                    cols = [c + w0 for c in plan_chunks(w1 - w0, chunk_size, True, ord_kind, perm_div)]
                    for wi, s in enumerate(starts):
                        for part in cols:
                            xt = x[s:s + win_size_t][:, :, :, part].permute(3, 1, 0, 2)
                            ct = concat_conds[s:s + win_size_t][:, :, :, part].permute(3, 1, 0, 2)
                            pred = cfg_noise(unets[r], xt, conds_t, t, ct, guidance_scale)
                            noises_t[s:s + win_size_t, :, :, part] = pred.permute(2, 1, 3, 0)
                        if s > 0:
                            noises_t[s:s + overlaps[wi - 1], :, :, w0:w1] *= np.sqrt(0.5)
            noises_t = adain(noises_t, noises)
            noises = (a ** 0.5) * noises_t + ((1 - a) ** 0.5) * noises
        x = sched.step(noises, t, x, generator=rng)[0]
        for u in unets:
            reset_oracle_pool(u)
    return x


# ---------------------------------------------------------------------------------------------
# DDIM inversion / reconstruction (reference invert.py:151-188, 215-244) as plain functions
# ---------------------------------------------------------------------------------------------
def ddim_coefs(sched, t, i, inversion: bool):
    ts = torch.flip(sched.timesteps, [0]) if inversion else sched.timesteps
    a_t = sched.alphas_cumprod[int(t)]
    if inversion:
        a_p = sched.alphas_cumprod[int(ts[i - 1])] if i > 0 else sched.final_alpha_cumprod
    else:
        a_p = sched.alphas_cumprod[int(ts[i + 1])] if i < len(ts) - 1 else sched.final_alpha_cumprod
    return a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5


def ddim_walk(unet, sched, x, conds, batch_size: int = 8, inversion: bool = True):
    """x_0 -> x_T (inversion) or x_T -> x_0: eps from the un-patched UNet without CFG, batch_size frames at a time."""
    ts = torch.flip(sched.timesteps, [0]) if inversion else sched.timesteps
    for i, t in enumerate(ts):
        eps = torch.cat([unet(x[b], t, encoder_hidden_states=conds[b]).sample
                         for b in torch.arange(len(x)).split(batch_size)])
        mu, sg, mu_p, sg_p = ddim_coefs(sched, t, i, inversion)
        if inversion:
            x = mu * ((x - sg_p * eps) / mu_p) + sg * eps
        else:
            x = mu_p * ((x - sg * eps) / mu) + sg_p * eps
    return x
