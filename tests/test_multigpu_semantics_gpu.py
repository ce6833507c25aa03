"""Multi-GPU semantics of path 1: each rank keeps its own VidToMe pool, i.e. the sharded run is
"the reference with the global-token pool reset at shard boundaries".  The oracle for that is
oracle/pipeline_ref.ddim_sample_oracle_sharded (== ddim_sample_oracle for world 1, checked on CPU in
tests/test_host_logic.py).  Here the B200 path is held against it:

  * on one device, two emulated ranks: two Generator/UNetB200 instances with set_shard(r, 2) run the REAL
    xy_pass / yt_pass (shard ranges, chunk plans, draw prefetch, per-rank pools and module generators) with
    per-rank host RNG streams; the all-reduce is the sum of the two partial tensors;
  * on a box with >= 2 GPUs, the same thing through torchrun + NCCL (tests/mgpu_worker.py), skipped otherwise.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UKW = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=768)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def make_inputs(dev, N=10, h=16, w=16):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 4, h, w, generator=g).repeat(N, 1, 1, 1).to(dev)
    base = torch.randn(1, 4, h, w, generator=g)
    cc = (0.18215 * (base + 0.1 * torch.randn(N, 4, h, w, generator=g))).to(dev)
    conds = torch.randn(2, 154, 768, generator=g).to(dev)
    conds_t = torch.randn(2, 77, 768, generator=g).to(dev)
    return x, cc, conds, conds_t


def oracle_sharded(dev, world, x, cc, conds, conds_t, n_timesteps, win):
    from oracle import pipeline_ref as P
    from oracle.unet_ref import make_unet

    unets = [make_unet(seed=0, **UKW).to(dev) for _ in range(world)]
    torch.manual_seed(12345)
    torch.cuda.manual_seed(12345)
    np.random.seed(12345)
    return P.ddim_sample_oracle_sharded(unets, x.clone(), conds, conds_t, cc, world, n_timesteps=n_timesteps, alpha_t=0.01,
                                        win_size_t=win, rng=[torch.Generator(device=dev).manual_seed(12345)] * len(x))


# measured on the B200: fp16 5.6e-3, bf16 1.15e-2 (and the pool-reset oracle differs from the single-pool one by 5.9e-3)
@pytest.mark.parametrize("adt,tol", [(torch.float16, 1.5e-2), (torch.bfloat16, 3e-2)])
def test_two_emulated_ranks_match_pool_reset_oracle(cuda, adt, tol):
    from oracle import pipeline_ref as P
    from oracle.unet_ref import make_unet
    from tclight_b200 import ops
    from tclight_b200.config_utils import default_config
    from tclight_b200.generate import Generator
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200
    from tclight_b200.unet import UNetB200

    world, n_timesteps, win = 2, 3, 6
    x, cc, conds, conds_t = make_inputs(cuda)
    N = len(x)
    want = oracle_sharded(cuda, world, x, cc, conds, conds_t, n_timesteps, win)
    single = oracle_sharded(cuda, 1, x, cc, conds, conds_t, n_timesteps, win)

    sd = {k: v.detach().cpu() for k, v in make_unet(seed=0, **UKW).state_dict().items()}
    torch.manual_seed(12345)
    torch.cuda.manual_seed(12345)
    np.random.seed(12345)
    gens = []
    for r in range(world):
        cfg = default_config(n_timesteps=n_timesteps, alpha_t=0.01, win_size_t=win)
        cfg.float_precision = "fp32"
        pipe = type("Pipe", (), {})()
        pipe.unet = UNetB200(sd, device=cuda, dtype=adt, block_out_channels=UKW["block_out_channels"])
        g = Generator(pipe, DPMSolverMultistepSchedulerB200(), cfg)
        g.set_shard(r, world)
        g._allreduce = lambda t: None            # the emulation sums the partial tensors below
        gens.append(g)
    rngs = [P.RankRng(12345) for _ in range(world)]
    sched = gens[0].scheduler
    sched.set_timesteps(n_timesteps, device=cuda)
    rng = [torch.Generator(device=cuda).manual_seed(12345)] * N
    xs = x.clone()
    ts = sched._timesteps_host
    for i in range(len(ts)):
        t = ts[i]
        parts, parts_t = [], []
        for r, g in enumerate(gens):
            with rngs[r]:
                parts.append(g.xy_pass(xs, conds.to(adt), t, cc, torch.zeros_like(xs)))
        noises = parts[0] + parts[1]
        f0, f1 = gens[0]._my_range(N)
        assert float(parts[0][f1:].abs().max()) == 0.0 and float(parts[1][:f1].abs().max()) == 0.0     # disjoint frame shards
        for r, g in enumerate(gens):
            with rngs[r]:
                parts_t.append(g.yt_pass(xs, conds_t.to(adt), t, cc, torch.zeros_like(xs)))
        noises_t = parts_t[0] + parts_t[1]
        alpha = 0.01 * 0.01 ** min(i / len(ts), 1)
        ops.adain_blend(noises_t, noises, alpha)
        xs = sched.step(noises, t, xs, generator=rng, return_dict=False)[0]
        for g in gens:
            g.post_iter(xs, t)
    err = rel_l2(xs, want)
    gap = rel_l2(single, want)
    print(f"2 emulated ranks {adt}: rel-L2 vs pool-reset oracle {err:.3e}; pool-reset vs single-pool oracle {gap:.3e}")
    assert err < tol


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_real_ranks_match_pool_reset_oracle(cuda, tmp_path):
    """torchrun x 2 over NCCL: path 1 against the pool-reset oracle, stage 1 / stage 2 data-parallel runs (peer-memory
    row shards) against the single-GPU run of the same library."""
    out = tmp_path / "mgpu.pt"
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    res = torch.load(out)
    x, cc, conds, conds_t = make_inputs(cuda)
    want = oracle_sharded(cuda, 2, x, cc, conds, conds_t, 3, 6)
    err = rel_l2(res["x"].to(cuda), want)
    print(f"2 NCCL ranks fp16: rel-L2 vs pool-reset oracle {err:.3e}")
    assert err < 1.5e-2
    assert res["stage2_loss_maxdiff"] < 1e-6 and res["stage1_loss_maxdiff"] < 1e-6
    assert res["stage2_img_meanabs"] < 5e-4
