"""Model construction for the B200 path: mirror of the reference's ``utils/model_utils.py`` (``init_iclight`` :12-94).

The reference assembles a diffusers ``StableDiffusionPipeline`` from realistic-vision-v51 and patches its UNet for
IC-Light: ``conv_in`` is widened to 8 input channels (latent 4 + condition 4, the new columns zero, :22-26), the
IC-Light offsets are ADDED to every UNet tensor (:50-54) and ``forward`` is hooked to concatenate the condition latent
(:35-43).  Here the same weight surgery is a pure function on state dicts (``iclight_merge_state_dict``) and the result
is loaded into ``UNetB200`` / ``AutoencoderKLB200`` (whose kernels do the concat while staging, tcl_stage_latent).

diffusers / safetensors / the checkpoints are not available offline; ``init_iclight`` therefore needs them at run time
and raises ``TclError`` when they are missing, and ``init_synthetic`` builds the same object graph from seeded random
weights (what bench.py, the tests and ``python -m tclight_b200.run --synthetic`` use).
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from ._lib import TclError
from .scheduler import DPMSolverMultistepSchedulerB200
from .unet import UNetB200
from .vae import AutoencoderKLB200
from .weights import random_state_dict, random_vae_state_dict


class DiffusionPipeline:
    """Name matters: the reference's apply_patch looks for a class called DiffusionPipeline (vidtome/patch.py:263)."""

    def __init__(self, unet=None, vae=None, text_encoder=None, tokenizer=None, scheduler=None):
        self.unet, self.vae, self.text_encoder, self.tokenizer, self.scheduler = unet, vae, text_encoder, tokenizer, scheduler


def iclight_merge_state_dict(sd_origin: Dict[str, torch.Tensor], sd_offset: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """model_utils.py:22-26, 50-54: zero-extend ``conv_in.weight`` from 4 to 8 input channels, then add the IC-Light
    offset to every tensor (strict: the key sets must match)."""
    sd = {k: v.clone() for k, v in sd_origin.items()}
    w = sd["conv_in.weight"]
    if w.shape[1] == 4:
        new_w = torch.zeros((w.shape[0], 8) + tuple(w.shape[2:]), dtype=w.dtype)
        new_w[:, :4] = w
        sd["conv_in.weight"] = new_w
    missing = set(sd) ^ set(sd_offset)
    if missing:
        raise TclError(f"IC-Light offset keys do not match the UNet ({len(missing)} differ, e.g. {sorted(missing)[:3]})")
    return {k: sd[k] + sd_offset[k].to(sd[k].dtype) for k in sd}


def _dtype(weight_dtype):
    if isinstance(weight_dtype, torch.dtype):
        return weight_dtype
    return {"fp16": torch.float16, "bf16": torch.bfloat16}.get(weight_dtype, torch.float16)


def init_iclight(device="cuda", model_path="./models/iclight_sd15_fc.safetensors", weight_dtype="fp16",
                 sd15_name="stablediffusionapi/realistic-vision-v51"):
    """Same signature and return value as the reference: (pipe, scheduler, 'iclight').  Needs diffusers, transformers,
    safetensors and the two checkpoints on disk (no download is attempted)."""
    try:
        import safetensors.torch as sf
        from diffusers import AutoencoderKL, UNet2DConditionModel
        from transformers import CLIPTextModel, CLIPTokenizer
    except Exception as e:  # noqa: BLE001
        raise TclError("init_iclight needs diffusers, transformers and safetensors (not installed here); use "
                       "init_synthetic() for seeded random weights") from e
    if not os.path.exists(model_path):
        raise TclError(f"IC-Light offsets not found at {model_path}")
    dt = _dtype(weight_dtype)
    tokenizer = CLIPTokenizer.from_pretrained(sd15_name, subfolder="tokenizer")
    text_encoder = CLIPTextModel.from_pretrained(sd15_name, subfolder="text_encoder").to(device=device, dtype=dt)
    vae_sd = AutoencoderKL.from_pretrained(sd15_name, subfolder="vae").state_dict()
    unet_sd = UNet2DConditionModel.from_pretrained(sd15_name, subfolder="unet").state_dict()
    merged = iclight_merge_state_dict(unet_sd, sf.load_file(model_path))
    unet = UNetB200(merged, device=device, dtype=dt)
    vae = AutoencoderKLB200(vae_sd, device=device, dtype=dt)
    scheduler = DPMSolverMultistepSchedulerB200()
    return DiffusionPipeline(unet, vae, text_encoder, tokenizer, scheduler), scheduler, "iclight"


def init_synthetic(device="cuda", weight_dtype="bf16", seed: int = 0, unet_channels=(320, 640, 1280, 1280),
                   vae_channels=(128, 256, 512, 512), cross_attention_dim: int = 768):
    """The same object graph from seeded random weights with diffusers' key layout (no text encoder: callers pass
    embeddings)."""
    dt = _dtype(weight_dtype)
    unet = UNetB200(random_state_dict(seed=seed, block_out_channels=unet_channels, cross_attention_dim=cross_attention_dim),
                    device=device, dtype=dt, block_out_channels=unet_channels)
    vae = AutoencoderKLB200(random_vae_state_dict(seed=seed + 1, block_out_channels=vae_channels), device=device, dtype=dt,
                            block_out_channels=vae_channels)
    scheduler = DPMSolverMultistepSchedulerB200()
    return DiffusionPipeline(unet, vae, None, None, scheduler), scheduler, "iclight"
