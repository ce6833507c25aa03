"""DDIM inversion on the B200 path: mirror of the reference's ``Inverter`` (invert.py:22-323; SURVEY.md
§8a row A17, §8f rank 3).  Same constructor / method names and argument meaning:

    Inverter(pipe, scheduler, config)(save_path, frame_ids=None)
    ddim_inversion(x, conds, save_path) -> x_T      ddim_sample(x, conds) -> x_0
    pred_noise(x, cond, t, batch_idx=None) -> eps   pred_next_x(x, eps, t, i, inversion=False) -> x'

The noise prediction is ``pipe.unet`` (UNetB200, no CFG, no VidToMe: the reference inverts before
``Generator.__init__`` patches the UNet, run.py:19-26), batched ``inversion.batch_size`` frames at a time;
the update is one fused kernel (tcl_ddim_next) over the whole latent instead of six tensor ops.

ControlNet / depth conditioning (invert.py:192-209) are outside SURVEY §8 and raise.  ``__call__`` needs a
VAE (``pipe.vae.encode``) and a data parser exactly like the reference; the latent-level methods above are
what the hot path consists of.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from ._lib import TclError, check, latent_code, lib, require_cuda, stream_ptr


class Inverter(nn.Module):
    def __init__(self, pipe, scheduler, config):
        super().__init__()
        self.device = config.device
        self.config = config
        inv = config.inversion
        if getattr(config, "sd_version", "1.5") == "depth":
            raise TclError("Inverter: the depth-conditioned UNet is outside the B200 hot path (SURVEY.md §8)")
        self.model_key = getattr(config, "model_key", None)
        fp = inv.float_precision if "float_precision" in inv else config.float_precision
        self.dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(fp, torch.float32)
        self.pipe = pipe
        self.unet = pipe.unet
        self.vae = getattr(pipe, "vae", None)
        self.tokenizer = getattr(pipe, "tokenizer", None)
        self.text_encoder = getattr(pipe, "text_encoder", None)
        self.control = inv.control
        if self.control != "none":
            raise TclError("Inverter: ControlNet-conditioned inversion is outside the B200 hot path (SURVEY.md §8)")
        scheduler.set_timesteps(inv.save_steps)
        self.timesteps_to_save = scheduler.timesteps
        scheduler.set_timesteps(inv.steps)
        self.scheduler = scheduler
        self.prompt = inv.prompt
        self.recon = inv.recon
        self.save_latents = inv.save_intermediate
        self.steps = inv.steps
        self.batch_size = inv.batch_size
        self.force = inv.force
        self.n_frames = inv.n_frames
        self.frame_height, self.frame_width = config.height, config.width
        self.work_dir = getattr(config, "work_dir", ".")
        self.data_parser = None           # set by the caller (video data parser, invert.py:93-96)

    # ---- hot path -----------------------------------------------------------------------------------
    @torch.no_grad()
    def pred_noise(self, x, cond, t, batch_idx=None):
        """invert.py:190-213 (plain SD branch)."""
        return self.unet(x, t, encoder_hidden_states=cond).sample

    def _coefs(self, t, i, inversion):
        """(mu, sigma, mu_prev, sigma_prev) as the reference computes them: fp32 0-dim tensor arithmetic
        (invert.py:219-237)."""
        sch = self.scheduler
        timesteps = reversed(sch.timesteps) if inversion else sch.timesteps
        ac = sch.alphas_cumprod
        a_t = ac[int(t)]
        if inversion:
            a_prev = ac[int(timesteps[i - 1])] if i > 0 else sch.final_alpha_cumprod
        else:
            a_prev = ac[int(timesteps[i + 1])] if i < len(timesteps) - 1 else sch.final_alpha_cumprod
        return (float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_prev ** 0.5), float((1 - a_prev) ** 0.5))

    @torch.no_grad()
    def pred_next_x(self, x, eps, t, i, inversion=False):
        require_cuda(x, eps)
        mu, sigma, mu_prev, sigma_prev = self._coefs(t, i, inversion)
        if inversion:
            args = (mu_prev, sigma_prev, mu, sigma)
        else:
            args = (mu, sigma, mu_prev, sigma_prev)
        x = x.contiguous()
        eps = eps.to(x.dtype).contiguous()
        out = torch.empty_like(x)
        check(lib.tcl_ddim_next(latent_code(x.dtype), eps.data_ptr(), x.data_ptr(), out.data_ptr(), x.numel(), *args,
                                stream_ptr()), "tcl_ddim_next")
        return out

    def _all_noise(self, x, conds, t):
        noises = torch.empty_like(x)
        for batch in torch.arange(len(x)).split(self.batch_size, dim=0):
            lo, hi = int(batch[0]), int(batch[-1]) + 1
            noises[lo:hi] = self.pred_noise(x[lo:hi], conds[lo:hi], t, batch_idx=batch)
        return noises

    @torch.no_grad()
    def ddim_inversion(self, x, conds, save_path=None):
        """invert.py:151-173: x_0 -> x_T over reversed(timesteps); the final latent is saved as
        ``noisy_latents_{t}.pt`` when ``save_path`` is given."""
        timesteps = reversed(self.scheduler.timesteps)
        t = None
        for i, t in enumerate(timesteps):
            noises = self._all_noise(x, conds, timesteps[i])
            x = self.pred_next_x(x, noises, t, i, inversion=True)
            if save_path is not None and self.save_latents and t in self.timesteps_to_save:
                torch.save(x, os.path.join(save_path, f"noisy_latents_{t}.pt"))
        if save_path is not None and t is not None:
            torch.save(x, os.path.join(save_path, f"noisy_latents_{t}.pt"))
        return x

    @torch.no_grad()
    def ddim_sample(self, x, conds):
        """invert.py:175-188: deterministic DDIM reconstruction."""
        timesteps = self.scheduler.timesteps
        for i, t in enumerate(timesteps):
            noises = self._all_noise(x, conds, t)
            x = self.pred_next_x(x, noises, t, i, inversion=False)
        return x

    # ---- surroundings (invert.py:104-149, 246-323) ------------------------------------------------------
    def check_latent_exists(self, save_path):
        save_timesteps = [self.scheduler.timesteps[0]]
        if self.save_latents:
            save_timesteps += list(self.timesteps_to_save)
        return all(os.path.exists(os.path.join(save_path, f"noisy_latents_{ts}.pt")) for ts in save_timesteps)

    @torch.no_grad()
    def prepare_cond(self, prompts, n_frames):
        if self.text_encoder is None or self.tokenizer is None:
            raise TclError("Inverter.prepare_cond needs pipe.tokenizer / pipe.text_encoder (CLIP is outside SURVEY §8)")
        if isinstance(prompts, str):
            prompts = [prompts] * n_frames
            cond = self.get_text_embeds(prompts[0])
            return torch.cat([cond] * n_frames), prompts
        return torch.cat([self.get_text_embeds(p) for p in prompts]), prompts

    @torch.no_grad()
    def get_text_embeds(self, prompt, negative_prompt=None, device="cuda"):
        ti = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                            return_tensors="pt")
        emb = self.text_encoder(ti.input_ids.to(device))[0]
        if negative_prompt is not None:
            ui = self.tokenizer(negative_prompt, padding="max_length", max_length=self.tokenizer.model_max_length,
                                return_tensors="pt")
            emb = torch.cat([self.text_encoder(ui.input_ids.to(device))[0], emb])
        return emb

    @torch.no_grad()
    def encode_imgs_batch(self, imgs):
        if self.vae is None:
            raise TclError("Inverter: pipe.vae is required to encode frames")
        out = []
        for img in imgs.split(self.batch_size, dim=0):
            out.append(self.vae.encode(2 * img.to(self.dtype) - 1).latent_dist.mean * 0.18215)
        return torch.cat(out)

    @torch.no_grad()
    def __call__(self, save_path, frame_ids=None):
        self.scheduler.set_timesteps(self.steps)
        # get_latents_dir (utils/VidToMe/utils.py:304-309, invert.py:276): one sub-directory per model key
        model_key = getattr(self, "model_key", None)
        save_path = os.path.join(save_path, model_key.split("/")[-1] if model_key is not None else "default")
        os.makedirs(save_path, exist_ok=True)
        if self.check_latent_exists(save_path) and not self.force:
            print(f"[INFO] inverted latents exist at: {save_path}. Skip inversion! Set 'inversion.force: True' to invert again.")
            return None
        if self.data_parser is None:
            raise TclError("Inverter.__call__: set .data_parser (object with load_video(frame_ids=...)) first")
        frames = self.data_parser.load_video(frame_ids=frame_ids)
        if isinstance(frames, (tuple, list)):
            frames = frames[0]
        if self.n_frames is not None:
            frames = frames[: self.n_frames]
        conds, prompts = self.prepare_cond(self.prompt, len(frames))
        with open(os.path.join(save_path, "inversion_prompts.txt"), "w") as f:
            f.write("\n".join(prompts))
        latents = self.encode_imgs_batch(frames)
        inverted = self.ddim_inversion(latents, conds, save_path)
        if self.recon:
            return inverted, self.ddim_sample(inverted, conds)
        return inverted
