"""The oracle restatements reproduce the golden vectors that oracle/make_goldens.py recorded from the
UNMODIFIED reference (its own compute_merge, Generator.ddim_sample/temporal_denoise/pred_noise,
exposure_align, unique_tensor_optimization) — runs anywhere, no reference tree needed."""
import copy
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_vidtome_compute_merge_golden():
    from oracle import make_goldens as G, vidtome_ref as V

    gold = torch.load(os.path.join(GOLD, "vidtome_compute_merge.pt"))
    state = V.MergeState(torch.Generator().manual_seed(7))
    for chunk, want in enumerate(gold):
        x = G.vidtome_inputs(chunk)
        merged, unmerge, _ = V.compute_merge(state, x, (8, 8), copy.deepcopy(G.VIDTOME_ARGS))
        assert torch.equal(merged, want["merged"])
        y = torch.arange(merged.numel(), dtype=torch.float32).reshape(merged.shape) / merged.numel()
        assert torch.equal(unmerge(y), want["unmerged"])
        assert torch.equal(state.global_tokens, want["pool"])


def test_sampler_golden():
    from oracle import make_goldens as G, pipeline_ref as P
    from oracle.unet_ref import make_unet

    gold = torch.load(os.path.join(GOLD, "sampler_ddim_multiaxis.pt"))
    x, cc, conds, conds_t = G.sampler_inputs()
    torch.manual_seed(12345)
    np.random.seed(12345)
    got = P.ddim_sample_oracle(make_unet(seed=0, **G.TINY_UNET), x.clone(), conds, conds_t, cc, n_timesteps=3, alpha_t=0.01,
                               win_size_t=6, rng=[torch.Generator().manual_seed(12345)] * len(x))
    # same torch build => bit-exact; allow fp32 noise for other BLAS builds
    assert torch.allclose(got, gold["x_final"], rtol=1e-4, atol=1e-4)


def test_scheduler_schedule_golden():
    from oracle.scheduler_ref import DPMSolverSDEKarras
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200

    gold = torch.load(os.path.join(GOLD, "sampler_ddim_multiaxis.pt"))
    for cls in (DPMSolverSDEKarras, DPMSolverMultistepSchedulerB200):
        s = cls()
        s.set_timesteps(3)
        assert torch.equal(s.timesteps, gold["timesteps"]) and torch.equal(s.sigmas, gold["sigmas"])


def test_postopt_golden():
    from oracle import postopt_ref as O

    gold = torch.load(os.path.join(GOLD, "postopt_stage12.pt"))
    for stage in (1, 2):
        edited, flows, masks, inv = O.synthetic_clip(n=5, h=176, w=184, seed=10 + stage)
        torch.manual_seed(20 + stage)
        batches = O.draw_batches(5, 4, 2)
        if stage == 1:
            img, _, losses = O.stage1_exposure(edited, flows, masks, batches)
        else:
            img, _, losses = O.stage2_uvt(edited, flows, masks, inv, batches)
        want = gold[f"stage{stage}"]
        got_l = torch.tensor(losses, dtype=torch.float64)
        assert got_l.shape == want["losses"].shape
        assert torch.allclose(got_l, want["losses"], rtol=0, atol=1e-6, equal_nan=True)
        assert torch.allclose(img.double().mean(dim=(2, 3)), want["image_mean"], atol=1e-6)
        assert torch.allclose(img[:, :, ::37, ::41], want["image_probe"], atol=1e-5)


def test_flowid_producer_golden():
    """soft masks / flow ids / unique inverse / bicubic warp restatements vs the reference's own outputs."""
    from oracle import flowid_ref as R

    gold = torch.load(os.path.join(GOLD, "flowid_producer.pt"))
    for seed, want in enumerate(gold):
        frames, fwd, bwd = R.synthetic_scene(n=6, h=40, w=56, seed=seed)
        masks = R.soft_mask_bwds(frames * 2 - 1, fwd, bwd, alpha=0.5)
        assert torch.allclose(masks, want["masks"], rtol=0, atol=2e-6)
        ids = R.flow_ids(frames, fwd, want["masks"], rgb_threshold=0.05)
        assert ids.dtype == torch.int32 and torch.equal(ids, want["ids"])
        assert torch.equal(R.unique_inverse(ids), want["inv"].reshape(-1))
        assert torch.allclose(R.warp(frames, bwd), want["warp"], rtol=0, atol=1e-6)
        # the scene exercises what it claims to
        assert 0.2 < (want["masks"] > 0.5).float().mean() < 0.95
        assert int(want["ids"].max()) + 1 < want["ids"].numel()


def test_ddim_inversion_golden():
    from oracle import make_goldens as G, pipeline_ref as P
    from oracle.scheduler_ref import DDIMRef
    from oracle.unet_ref import make_unet
    from tclight_b200.scheduler import DDIMSchedulerB200

    gold = torch.load(os.path.join(GOLD, "ddim_inversion.pt"))
    x, conds = G.inversion_inputs()
    sch = DDIMRef()
    sch.set_timesteps(5)
    assert torch.equal(sch.timesteps, gold["timesteps"])
    prod = DDIMSchedulerB200()
    prod.set_timesteps(5)
    assert torch.equal(prod.timesteps, gold["timesteps"]) and torch.equal(prod.alphas_cumprod, sch.alphas_cumprod)
    assert torch.equal(prod.final_alpha_cumprod, sch.final_alpha_cumprod)
    unet = make_unet(seed=0, **G.INV_UNET)
    with torch.no_grad():
        xT = P.ddim_walk(unet, sch, x, conds, 4, True)
        x0 = P.ddim_walk(unet, sch, xT, conds, 4, False)
    assert torch.allclose(xT, gold["x_T"], rtol=1e-4, atol=1e-4)
    assert torch.allclose(x0, gold["x_recon"], rtol=1e-4, atol=1e-4)
