"""TEST INFRASTRUCTURE — builds an instance of the reference's own ``generate.Generator`` (unmodified
code from /root/reference, imported through oracle/refshim.py) around the restated UNet and
scheduler, so that the reference's ``ddim_sample`` / ``temporal_denoise`` / ``pred_noise`` /
``exposure_align`` / ``unique_tensor_optimization`` run on CPU.  Build-container only.
"""
from __future__ import annotations

import types

import torch

from . import refshim
from .scheduler_ref import DPMSolverSDEKarras
from .unet_ref import make_unet

DEFAULT_GEN = dict(
    guidance_scale=2.0, n_timesteps=25, chunk_size=4, chunk_ord="mix-4", local_merge_ratio=0.6,
    merge_global=True, global_merge_ratio=0.5, global_rand=0.5, align_batch=True, max_downsample=2,
    noise_mode="same", alpha_t=0.0, final_factor_t=0.01, win_size_t=64, seed=12345,
)
DEFAULT_OPT = dict(
    lambda_dssim=0.2, lambda_flow=0.8, lambda_tv=0.05, epochs_exposure=35, epochs=70, batch_size=16,
    feature_lr=0.05, exposure_lr_init=0.01, exposure_lr_final=0.001, exposure_lr_delay_steps=0,
    exposure_lr_delay_mult=0.0,
)


class DiffusionPipeline:
    """Name matters: reference patch.py:263 checks ``isinstance_str(model, "DiffusionPipeline")``."""


class _Pipe(DiffusionPipeline):
    """Minimal stand-in for StableDiffusionPipeline: ``apply_patch`` only needs ``.unet``."""

    def __init__(self, unet):
        self.unet = unet


def make_reference_generator(unet=None, device="cpu", dtype=torch.float32, gen=None, opt=None, unet_kw=None):
    """Returns (generator, ref_namespace).  ``generator`` is a real reference ``Generator`` whose
    __init__ (which needs diffusers pipelines / data parsers) is bypassed; the attributes its
    hot-path methods read are injected with the values the reference __init__ would set
    (generate.py:50-78, generate_utils.py:21-96)."""
    ref = refshim.import_reference()
    G = ref.generate.Generator
    g = object.__new__(G)
    torch.nn.Module.__init__(g)
    cfg = dict(DEFAULT_GEN)
    cfg.update(gen or {})
    oc = dict(DEFAULT_OPT)
    oc.update(opt or {})
    if unet is None:
        unet = make_unet(seed=0, dtype=dtype, **(unet_kw or {})).to(device)
    sched = DPMSolverSDEKarras()
    sched.set_timesteps(cfg["n_timesteps"], device=device)
    g.device = device
    g.dtype = dtype
    g.seed = cfg["seed"]
    g.model_key = "iclight"
    g.pipe = _Pipe(unet)
    g.unet = unet
    g.scheduler = sched
    g.n_timesteps = cfg["n_timesteps"]
    g.batch_size = 2
    g.use_pnp = False
    g.use_depth = False
    g.use_controlnet = False
    g.control = "none"
    for k in ("chunk_size", "merge_global", "local_merge_ratio", "global_merge_ratio", "global_rand",
              "align_batch", "guidance_scale", "noise_mode", "max_downsample", "win_size_t", "alpha_t",
              "final_factor_t"):
        setattr(g, k, cfg[k])
    g.chunk_ord = cfg["chunk_ord"]
    if "mix" in g.chunk_ord:  # generate_utils.py:88-91
        g.perm_div = float(g.chunk_ord.split("-")[-1]) if "-" in g.chunk_ord else 3.0
        g.chunk_ord = "mix"
    # post-opt (generate.py:50-64)
    g.apply_opt = True
    g.lambda_dssim, g.lambda_flow, g.lambda_tv = oc["lambda_dssim"], oc["lambda_flow"], oc["lambda_tv"]
    g.epochs_exposure, g.epochs, g.opt_batch_size = oc["epochs_exposure"], oc["epochs"], oc["batch_size"]
    g.feature_lr = oc["feature_lr"]
    g.exposure_lr_init, g.exposure_lr_final = oc["exposure_lr_init"], oc["exposure_lr_final"]
    g.exposure_lr_delay_steps, g.exposure_lr_delay_mult = oc["exposure_lr_delay_steps"], oc["exposure_lr_delay_mult"]
    g.data_parser = types.SimpleNamespace(unq_inv=None)
    g.dataset = None
    # generate_utils.py:98-100
    ref.patch.apply_patch(g.pipe, g.local_merge_ratio, g.merge_global, g.global_merge_ratio,
                          seed=g.seed, batch_size=g.batch_size, align_batch=g.align_batch,
                          global_rand=g.global_rand)
    return g, ref


def seed_everything(seed: int):
    """reference utils/VidToMe/utils.py:70-74."""
    import random

    import numpy as np

    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
