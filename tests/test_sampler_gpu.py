"""Path-1 sampler parity on the GPU: the B200 Generator mirror (CUDA kernels) vs the oracle
pipeline — the restated reference loop (oracle/pipeline_ref.py, pinned to the reference's own
generate.py in the build container) over the oracle UNet/scheduler, same seeds, same device.

Tolerances: scheduler / AdaIN kernels reproduce the reference's 16-bit roundings => compared at
<= 2 ulp-level (atol 2e-3 relative to unit-scale latents); full sampler (fp16 UNet vs fp32 oracle
UNet over several steps with VidToMe) => bounds <= 3x the measured error, per test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("ldt", [torch.float16, torch.float32])
def test_scheduler_step_matches_oracle(cuda, ldt):
    from oracle.scheduler_ref import DPMSolverSDEKarras
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200

    N, h, w = 5, 12, 20
    ref, mine = DPMSolverSDEKarras(), DPMSolverMultistepSchedulerB200()
    ref.set_timesteps(25, device=cuda)
    mine.set_timesteps(25, device=cuda)
    assert torch.equal(ref.timesteps, mine.timesteps) and torch.equal(ref.sigmas, mine.sigmas)
    g1 = [torch.Generator(device=cuda).manual_seed(5)] * N
    g2 = [torch.Generator(device=cuda).manual_seed(5)] * N
    torch.manual_seed(0)
    x1 = torch.randn(N, 4, h, w, device=cuda).to(ldt)
    x2 = x1.clone()
    for i, t in enumerate(ref.timesteps):
        eps = torch.randn(N, 4, h, w, device=cuda).to(ldt)
        x1 = ref.step(eps, t, x1, generator=g1)[0]
        x2 = mine.step(eps, t, x2, generator=g2)[0]
        if ldt == torch.float16:
            # identical op-by-op roundings: allow 1 fp16 ulp for FMA-contraction differences
            assert (x1.float() - x2.float()).abs().max().item() <= 2 * torch.finfo(torch.float16).eps * x1.float().abs().max().item(), i
        else:
            assert torch.allclose(x1, x2, rtol=1e-5, atol=1e-5), i
        x2 = x1.clone()      # keep trajectories locked so each step is tested in isolation
        mine.model_outputs[1] = ref.model_outputs[1].clone()


@pytest.mark.parametrize("ldt", [torch.float16, torch.float32])
def test_adain_blend_matches_reference_ops(cuda, ldt):
    from tclight_b200 import ops

    torch.manual_seed(1)
    N, h, w = 6, 23, 40
    nt = (torch.randn(N, 4, h, w, device=cuda) * 1.7 + 0.3).to(ldt)
    nz = (torch.randn(N, 4, h, w, device=cuda) * 0.8 - 0.1).to(ldt)
    alpha = 0.01 * 0.01 ** (3 / 25)

    def mean_std(f, eps=1e-5):     # utils/general_utils.py:137-146
        n, c = f.shape[:2]
        var = f.view(n, c, -1).var(dim=2) + eps
        return f.view(n, c, -1).mean(dim=2).view(n, c, 1, 1), var.sqrt().view(n, c, 1, 1)

    sm, ss = mean_std(nz)
    cm, cs = mean_std(nt)
    ad = (nt - cm) / cs * ss + sm
    bl = (alpha ** 0.5) * ad + ((1 - alpha) ** 0.5) * nz
    a2, b2 = nt.clone(), nz.clone()
    ops.adain_blend(a2, b2, alpha)
    tol = 4e-3 if ldt == torch.float16 else 1e-5
    assert (a2.float() - ad.float()).abs().max().item() <= tol * ad.float().abs().max().item()
    assert (b2.float() - bl.float()).abs().max().item() <= tol * bl.float().abs().max().item()


def _seed_all():
    torch.manual_seed(12345)
    torch.cuda.manual_seed(12345)
    np.random.seed(12345)


def _sampler_pair(cuda, adt, ukw, **gen_kw):
    """(oracle UNet fp32, B200 Generator in activation dtype `adt`) with identical seeded weights."""
    from oracle.unet_ref import make_unet
    from tclight_b200.config_utils import default_config
    from tclight_b200.generate import Generator
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200
    from tclight_b200.unet import UNetB200

    ref_unet = make_unet(seed=0, **ukw).to(cuda)
    sd = {k: v.detach().cpu() for k, v in ref_unet.state_dict().items()}
    mine_unet = UNetB200(sd, device=cuda, dtype=adt, block_out_channels=ukw.get("block_out_channels", (320, 640, 1280, 1280)))
    cfg = default_config(**gen_kw)
    cfg.float_precision = "fp32"       # latents in fp32 on both sides; UNet activations 16-bit vs the fp32 oracle
    pipe = type("Pipe", (), {})()
    pipe.unet = mine_unet
    return ref_unet, Generator(pipe, DPMSolverMultistepSchedulerB200(), cfg)


def _inputs(cuda, N, h, w, seed=7):
    torch.manual_seed(seed)
    x = torch.randn(1, 4, h, w, device=cuda).repeat(N, 1, 1, 1)
    base = torch.randn(1, 4, h, w, device=cuda)
    cc = 0.18215 * (base + 0.1 * torch.randn(N, 4, h, w, device=cuda))
    conds = torch.randn(2, 154, 768, device=cuda)
    conds_t = torch.randn(2, 77, 768, device=cuda)
    return x, cc, conds, conds_t


# Loop-level bounds are <= 3x the error measured on the B200 (recorded next to each bound); the discrete VidToMe
# decisions make the error of a multi-step run heavy-tailed, so the measured value is printed for the record.
@pytest.mark.parametrize("adt,tol", [(torch.float16, 1.5e-2), (torch.bfloat16, 3e-2)])
def test_ddim_sample_multi_axis_vs_oracle(cuda, adt, tol):
    """4 steps, 8 frames, multi-axis, VidToMe on, tiny widths.  Measured on the B200: fp16 5.0e-3, bf16 9.9e-3."""
    from oracle import pipeline_ref as P

    ref_unet, gen = _sampler_pair(cuda, adt, dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=768),
                                  n_timesteps=4, alpha_t=0.01, win_size_t=6)
    N, h, w = 8, 16, 16
    x, cc, conds, conds_t = _inputs(cuda, N, h, w)
    _seed_all()
    gen.rng = [torch.Generator(device=cuda).manual_seed(12345)] * N
    got = gen.ddim_sample(x.clone(), conds.to(adt), conds_t.to(adt), cc)
    _seed_all()
    want = P.ddim_sample_oracle(ref_unet, x.clone(), conds, conds_t, cc, n_timesteps=4, alpha_t=0.01, win_size_t=6,
                                rng=[torch.Generator(device=cuda).manual_seed(12345)] * N)
    err = rel_l2(got, want)
    print(f"multi-axis ddim_sample {adt} (4 steps, 8 frames, VidToMe on): rel-L2 {err:.3e}")
    assert err < tol


@pytest.mark.parametrize("adt,tol", [(torch.float16, 1.5e-2), (torch.bfloat16, 2.5e-2)])
def test_baseline_config1_sd15_widths(cuda, adt, tol):
    """Measured on the B200: fp16 5.3e-3, bf16 8.6e-3.  BASELINE.json configs[0] exactly: 8 frames, 256x256 (latent 32x32), 4 denoising steps, single axis,
    **SD-1.5 widths** (320/640/1280/1280: ds-1 merging at head dim 40, ds-2 at head dim 80), VidToMe on, L = 154 —
    B200 path vs the oracle loop on the same seeds."""
    from oracle import pipeline_ref as P

    ref_unet, gen = _sampler_pair(cuda, adt, {}, n_timesteps=4, alpha_t=0.0)
    N, h, w = 8, 32, 32
    x, cc, conds, conds_t = _inputs(cuda, N, h, w, seed=11)
    _seed_all()
    gen.rng = [torch.Generator(device=cuda).manual_seed(12345)] * N
    got = gen.ddim_sample(x.clone(), conds.to(adt), conds_t.to(adt), cc)
    _seed_all()
    want = P.ddim_sample_oracle(ref_unet, x.clone(), conds, conds_t, cc, n_timesteps=4, alpha_t=0.0,
                                rng=[torch.Generator(device=cuda).manual_seed(12345)] * N)
    err = rel_l2(got, want)
    print(f"config 1 (8f 256x256, 4 steps, single axis, SD-1.5 widths) {adt}: rel-L2 {err:.3e}")
    assert err < tol


@pytest.mark.parametrize("adt,tol", [(torch.float16, 1.5e-2), (torch.bfloat16, 3e-2)])
def test_multi_axis_step_sd15_widths(cuda, adt, tol):
    """Measured on the B200: fp16 5.7e-3, bf16 9.9e-3.  Two full multi-axis steps (xy pass + yt pass over 2 overlapping windows + AdaIN/blend + DPM-Solver++ update) at
    SD-1.5 widths: the yt 'images' are (frames x height) = 6x32 per latent column."""
    from oracle import pipeline_ref as P

    ref_unet, gen = _sampler_pair(cuda, adt, {}, n_timesteps=2, alpha_t=0.01, win_size_t=6)
    N, h, w = 8, 32, 32
    x, cc, conds, conds_t = _inputs(cuda, N, h, w, seed=13)
    _seed_all()
    gen.rng = [torch.Generator(device=cuda).manual_seed(12345)] * N
    got = gen.ddim_sample(x.clone(), conds.to(adt), conds_t.to(adt), cc)
    _seed_all()
    want = P.ddim_sample_oracle(ref_unet, x.clone(), conds, conds_t, cc, n_timesteps=2, alpha_t=0.01, win_size_t=6,
                                rng=[torch.Generator(device=cuda).manual_seed(12345)] * N)
    err = rel_l2(got, want)
    print(f"multi-axis, SD-1.5 widths, 2 steps {adt}: rel-L2 {err:.3e}")
    assert err < tol
