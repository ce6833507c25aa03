"""Multi-GPU host logic on CPU: 2 processes over gloo.  Each rank computes its frame shard (xy pass)
and column shard (yt pass) with a stand-in noise predictor, the partial tensors are summed with
all-reduce, and the result must equal the single-process result."""
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_pred(self, x, cond, t, concat_conds=None, batch_idx=None, sl_i=None, out=None):
    out.copy_(x * 2.0 + (concat_conds if concat_conds is not None else 0) + float(t) * 1e-3)
    return out


def _make(rank, world):
    from tclight_b200.generate import Generator

    g = types.SimpleNamespace(chunk_size=4, merge_global=True, chunk_ord="mix", perm_div=4.0, win_size_t=6, guidance_scale=2.0,
                              pipe=types.SimpleNamespace(unet=torch.nn.Identity()))     # un-patched stand-in: no VidToMe draws to prefetch
    for name in ("get_chunks", "_my_range", "_allreduce", "set_shard", "xy_pass", "yt_pass"):
        setattr(g, name, types.MethodType(getattr(Generator, name), g))
    g.temporal_windows = Generator.temporal_windows
    g.pred_noise = types.MethodType(_fake_pred, g)
    g.set_shard(rank, world)
    return g


def _inputs():
    gen = torch.Generator().manual_seed(0)
    return torch.randn(11, 4, 6, 10, generator=gen), torch.randn(11, 4, 6, 10, generator=gen)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, cc = _inputs()
    g = _make(rank, world)
    np.random.seed(rank); torch.manual_seed(rank)
    noises = torch.zeros_like(x)
    g.xy_pass(x, None, 801, cc, noises)
    noises_t = torch.zeros_like(x)
    g.yt_pass(x, None, 801, cc, noises_t, scale=lambda t, s: t.mul_(s))
    if rank == 0:
        q.put((noises, noises_t))
    dist.destroy_process_group()


def test_sharded_passes_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, cc = _inputs()
    g = _make(0, 1)
    want = torch.zeros_like(x)
    g.xy_pass(x, None, 801, cc, want)
    want_t = torch.zeros_like(x)
    g.yt_pass(x, None, 801, cc, want_t, scale=lambda t, s: t.mul_(s))
    assert torch.allclose(got, want) and torch.allclose(got_t, want_t)
    # shard ranges tile the axis exactly
    for n in (1, 7, 160, 300):
        for w in (1, 2, 4, 8):
            edges = [_make(r, w)._my_range(n) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


# ---- data-parallel optimiser host logic (tclight_b200/postopt.py): batch slicing and the CPU-RNG hand-over ----
def test_shard_batch_partitions_sorted_batch():
    from tclight_b200.postopt import shard_batch

    idxs = [17, 3, 250, 0, 42, 299, 8, 120, 64, 5, 199, 77]
    for world in (1, 2, 3, 8):
        parts = [shard_batch(idxs, r, world) for r in range(world)]
        got = [i for p, _, _ in parts for i in p]
        assert got == sorted(idxs)                                  # contiguous slices of the sorted batch, nothing lost
        assert all(n == len(idxs) and nv == len(idxs) - 1 for _, n, nv in parts)     # global normalisers on every rank
        sizes = [len(p) for p, _, _ in parts]
        assert max(sizes) - min(sizes) <= 1


def _rng_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tclight_b200.postopt import _sync_cpu_rng, batch_iterator

    torch.manual_seed(5)
    if rank > 0:
        torch.randperm(2 + rank)            # the sharded denoising passes leave the ranks' CPU RNGs in different states
    _sync_cpu_rng(torch.device("cpu"))
    draws = [[int(i) for i in b] for b in batch_iterator(7, 4)]
    q.put((rank, draws))
    dist.destroy_process_group()


def test_dp_ranks_draw_identical_batches_after_rng_sync():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rng_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(5)
    from tclight_b200.postopt import batch_iterator
    want = [[int(i) for i in b] for b in batch_iterator(7, 4)]
    assert out[0] == out[1] == want         # = what rank 0 / a single-GPU run draws
