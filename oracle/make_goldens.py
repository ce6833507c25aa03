"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference through oracle/refshim.py) on seeded inputs.  Run in the build container:

    python -m oracle.make_goldens

The goldens pin the oracle restatements on machines where the reference tree does not exist (the
GPU box): tests/test_golden.py replays the same seeded inputs through oracle/* and compares.
All tensors are tiny (a few hundred KB in total).
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from . import harness, refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

VIDTOME_ARGS = dict(max_downsample=2, generator=None, seed=123, batch_size=2, align_batch=True, merge_global=True,
                    global_merge_ratio=0.5, local_merge_ratio=0.6, global_rand=0.5, target_stride=4)
TINY_UNET = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)


class _Mod:
    pass


def vidtome_inputs(chunk: int, F: int = 4, n: int = 64, C: int = 32):
    g = torch.Generator().manual_seed(100 + chunk)
    return torch.randn(2 * F, n, C, generator=g)


def golden_vidtome(ref):
    mod = _Mod()
    mod.generator = torch.Generator().manual_seed(7)
    info = dict(size=(8, 8), args=copy.deepcopy(VIDTOME_ARGS))
    out = []
    for chunk in range(3):
        x = vidtome_inputs(chunk)
        m, u, merged = ref.patch.compute_merge(mod, x, info)
        y = torch.arange(merged.numel(), dtype=torch.float32).reshape(merged.shape) / merged.numel()
        out.append(dict(merged=merged.clone(), unmerged=u(y).clone(), pool=mod.global_tokens.clone()))
    return out


def sampler_inputs():
    g = torch.Generator().manual_seed(3)
    N, h, w = 8, 16, 16
    x = torch.randn(1, 4, h, w, generator=g).repeat(N, 1, 1, 1)
    cc = torch.randn(N, 4, h, w, generator=g) * 0.18215
    conds = torch.randn(2, 20, 64, generator=g)
    conds_t = torch.randn(2, 10, 64, generator=g)
    return x, cc, conds, conds_t


def golden_sampler(ref):
    from .unet_ref import make_unet

    x, cc, conds, conds_t = sampler_inputs()
    g, _ = harness.make_reference_generator(unet=make_unet(seed=0, **TINY_UNET), gen=dict(n_timesteps=3, alpha_t=0.01, win_size_t=6))
    torch.manual_seed(12345)
    np.random.seed(12345)
    g.rng = [torch.Generator().manual_seed(12345)] * len(x)
    out = g.ddim_sample(x.clone(), conds, conds_t, cc)
    return dict(x_final=out.clone(), timesteps=g.scheduler.timesteps.clone(), sigmas=g.scheduler.sigmas.clone())


def golden_postopt(ref):
    from . import postopt_ref as O
    from .unet_ref import make_unet

    res = {}
    for stage in (1, 2):
        edited, flows, masks, inv = O.synthetic_clip(n=5, h=176, w=184, seed=10 + stage)
        opt = dict(epochs=2, epochs_exposure=2, batch_size=4)
        g, _ = harness.make_reference_generator(unet=make_unet(seed=0, block_out_channels=(64, 64, 64, 64), cross_attention_dim=64), opt=opt)
        g.dataset = ref.dataloader.OptDataset(edited.clone(), flows.clone(), masks.clone(), device="cpu")
        g.data_parser.unq_inv = inv.clone()
        torch.manual_seed(20 + stage)
        if stage == 1:
            real_eye = torch.eye
            ref.generate.torch.eye = lambda *a, device=None, **k: real_eye(*a, **k)   # generate.py:378 hard-codes "cuda"
            try:
                img, losses = g.exposure_align()
            finally:
                ref.generate.torch.eye = real_eye
        else:
            img, losses = g.unique_tensor_optimization()
        res[f"stage{stage}"] = dict(losses=torch.tensor(losses, dtype=torch.float64), image_mean=img.double().mean(dim=(2, 3)).clone(),
                                    image_probe=img[:, :, ::37, ::41].clone())
    return res


def golden_flowid(ref):
    """reference get_soft_mask_bwds / get_flowid / voxelization / warp_flow on the seeded scene of
    oracle/flowid_ref.synthetic_scene (sub-pixel flows, occluder, collisions, out-of-frame targets)."""
    from . import flowid_ref as R

    out = []
    for seed in (0, 1):
        frames, fwd, bwd = R.synthetic_scene(n=6, h=40, w=56, seed=seed)
        masks = ref.flow_utils.get_soft_mask_bwds(frames * 2 - 1, fwd, bwd, alpha=0.5)
        ids = ref.flow_utils.get_flowid(frames, fwd, masks, rgb_threshold=0.05)
        inv = ref.general_utils.voxelization(ids.view(-1, 1), frames.permute(0, 2, 3, 1).reshape(-1, 3), None, None)
        out.append(dict(masks=masks.clone(), ids=ids.clone(), inv=inv.clone(), warp=ref.flow_utils.warp_flow(frames, bwd).clone()))
    return out


INV_UNET = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64, in_channels=4)


def inversion_inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6, 4, 16, 16, generator=g)
    conds = torch.randn(1, 10, 64, generator=g).repeat(6, 1, 1)
    return x, conds


def golden_inversion():
    """the reference's own Inverter.ddim_inversion / ddim_sample (invert.py:151-188) around the oracle UNet
    (un-patched, no CFG) and the restated DDIM schedule, 5 steps, batch 4."""
    import importlib
    import tempfile

    from .scheduler_ref import DDIMRef
    from .unet_ref import make_unet

    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        inv = importlib.import_module("invert")
    finally:
        os.chdir(cwd)
    I = object.__new__(inv.Inverter)
    torch.nn.Module.__init__(I)
    I.device, I.dtype, I.unet, I.scheduler = "cpu", torch.float32, make_unet(seed=0, **INV_UNET), DDIMRef()
    I.scheduler.set_timesteps(5)
    I.batch_size, I.use_depth, I.control, I.save_latents = 4, False, "none", False
    x, conds = inversion_inputs()
    xT = I.ddim_inversion(x, conds, tempfile.mkdtemp())
    x0 = I.ddim_sample(xT, conds)
    return dict(x_T=xT.clone(), x_recon=x0.clone(), timesteps=I.scheduler.timesteps.clone())


def main():
    ref = refshim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.save(golden_flowid(ref), os.path.join(OUT, "flowid_producer.pt"))
    torch.save(golden_inversion(), os.path.join(OUT, "ddim_inversion.pt"))
    torch.save(golden_vidtome(ref), os.path.join(OUT, "vidtome_compute_merge.pt"))
    torch.save(golden_sampler(ref), os.path.join(OUT, "sampler_ddim_multiaxis.pt"))
    torch.save(golden_postopt(ref), os.path.join(OUT, "postopt_stage12.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
