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
