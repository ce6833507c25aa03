"""B200 mirror of the reference pipeline driver: ``Generator`` of generate.py:41-630 on top of
``VidToMeGenerator`` (utils/VidToMe/generate_utils.py:20-238).  Same constructor, method names,
argument meaning and RNG call pattern for the two hot paths:

  * path 1: ``ddim_sample`` (:207-239), ``temporal_denoise`` (:241-284), ``pred_noise`` (:287-352),
    ``get_chunks`` (generate_utils.py:174-205), ``pre_iter/post_iter`` (:228-238);
  * path 2: ``exposure_align`` (:354-451), ``unique_tensor_optimization`` (:453-533)
    (implemented in tclight_b200/postopt.py and bound here).

Everything tensor-sized runs in libtclight.so.  Host-side differences that do not change
results: chunks are passed to the UNet as strided *views* of the latent (no gather copies), the
CFG combine writes straight into ``noises[chunk]``, and timesteps are host ints (no per-call
device sync).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops, vidtome
from ._lib import TclError


def _as_range(idx: torch.Tensor):
    """A chunk produced by get_chunks is a run of consecutive indices: return (lo, hi)."""
    lo, hi = int(idx[0]), int(idx[-1]) + 1
    if hi - lo != len(idx):
        raise TclError("chunk indices must be consecutive")
    return lo, hi


class VidToMeGenerator(nn.Module):
    """reference utils/VidToMe/generate_utils.py:20-96 (constructor wiring)."""

    def __init__(self, pipe, scheduler, config):
        super().__init__()
        self.device = config.device
        self.seed = config.seed
        self.model_key = config.model_key
        self.config = config
        gene = config.generation
        float_precision = gene.float_precision if "float_precision" in gene else config.float_precision
        self.dtype = torch.float16 if float_precision == "fp16" else (torch.bfloat16 if float_precision == "bf16" else torch.float32)
        self.pipe = pipe
        self.unet = pipe.unet
        self.vae = getattr(pipe, "vae", None)
        self.tokenizer = getattr(pipe, "tokenizer", None)
        self.text_encoder = getattr(pipe, "text_encoder", None)
        self.n_timesteps = gene.n_timesteps
        scheduler.set_timesteps(gene.n_timesteps, device=self.device)
        self.scheduler = scheduler
        self.batch_size = 2
        self.control = gene.control
        if self.control not in ("none", None) or config.sd_version == "depth":
            raise NotImplementedError("PnP / ControlNet / depth conditioning are not involved in TC-Light "
                                      "(configs/tclight_default.yaml:12,29) and are not implemented")
        self.use_depth = self.use_controlnet = self.use_pnp = False
        self.chunk_size = gene.chunk_size
        self.chunk_ord = gene.chunk_ord
        self.merge_global = gene.merge_global
        self.local_merge_ratio = gene.local_merge_ratio
        self.global_merge_ratio = gene.global_merge_ratio
        self.global_rand = gene.global_rand
        self.align_batch = gene.align_batch
        self.prompt = gene.prompt
        self.negative_prompt = gene.negative_prompt
        self.guidance_scale = gene.guidance_scale
        self.save_frame = gene.save_frame
        self.work_dir = config.work_dir
        if "mix" in self.chunk_ord:   # generate_utils.py:88-91
            self.perm_div = float(self.chunk_ord.split("-")[-1]) if "-" in self.chunk_ord else 3.0
            self.chunk_ord = "mix"
        self.activate_vidtome()

    def activate_vidtome(self):
        # generate_utils.py:98-100 (max_downsample from the YAML is stored but never passed)
        vidtome.apply_patch(self.pipe, self.local_merge_ratio, self.merge_global, self.global_merge_ratio,
                            seed=self.seed, batch_size=self.batch_size,
                            align_batch=self.use_pnp or self.align_batch, global_rand=self.global_rand)

    def get_chunks(self, flen):
        """generate_utils.py:174-205 — same CPU RNG calls in the same order."""
        x_index = torch.arange(flen)
        rand_first = np.random.randint(0, self.chunk_size) + 1
        chunks = x_index[rand_first:].split(self.chunk_size, dim=0)
        chunks = [x_index[:rand_first]] + list(chunks) if len(chunks[0]) > 0 else [x_index[:rand_first]]
        if np.random.rand() > 0.5:
            chunks = chunks[::-1]
        if self.merge_global is False:
            return chunks
        if self.chunk_ord == "rand":
            order = torch.randperm(len(chunks))
        elif self.chunk_ord == "mix":
            randord = torch.randperm(len(chunks)).tolist()
            rand_len = int(len(randord) / self.perm_div)
            seqord = sorted(randord[rand_len:])
            if rand_len > 0:
                randord = randord[:rand_len]
                if abs(seqord[-1] - randord[-1]) < abs(seqord[0] - randord[-1]):
                    seqord = seqord[::-1]
                order = randord + seqord
            else:
                order = seqord
        else:
            order = torch.arange(len(chunks))
        return [chunks[i] for i in order]

    def pre_iter(self, x, t):
        return None  # only PnP does work here (generate_utils.py:228-233)

    def post_iter(self, x, t):
        vidtome.assert_draws_consumed(self.pipe)
        if self.merge_global:
            vidtome.update_patch(self.pipe, global_tokens=None)   # generate_utils.py:235-238


class Generator(VidToMeGenerator):
    def __init__(self, pipe, scheduler, config):
        super().__init__(pipe, scheduler, config)
        self.config = config
        self._init_generation(config.generation)
        self._init_post_optimization(config.post_opt)
        self.dataset = None
        self.data_parser = getattr(pipe, "data_parser", None)
        self.rng: Optional[List[torch.Generator]] = None

    def _init_post_optimization(self, c):   # generate.py:50-64
        self.apply_opt = c.apply_opt
        self.lambda_dssim, self.lambda_flow, self.lambda_tv = c.lambda_dssim, c.lambda_flow, c.lambda_tv
        self.epochs_exposure, self.epochs, self.opt_batch_size = c.epochs_exposure, c.epochs, c.batch_size
        self.feature_lr = c.feature_lr
        self.exposure_lr_init, self.exposure_lr_final = c.exposure_lr_init, c.exposure_lr_final
        self.exposure_lr_delay_steps, self.exposure_lr_delay_mult = c.exposure_lr_delay_steps, c.exposure_lr_delay_mult

    def _init_generation(self, g):          # generate.py:66-78
        self.background_cond = g.background_cond
        self.noise_mode = g.noise_mode
        self.max_downsample = g.max_downsample
        self.win_size_t = g.win_size_t
        self.alpha_t = g.alpha_t
        self.final_factor_t = g.final_factor_t
        self.prompt_t = g.prompt_t
        self.negative_prompt_t = g.negative_prompt_t

    # ---------------------------------------------------------------- multi-GPU sharding
    def set_shard(self, rank: int, world: int):
        """One process per GPU: the xy pass is sharded over contiguous frame ranges,
        the yt pass over contiguous latent-column ranges; the per-rank partial noise tensors are
        summed with one NCCL all-reduce each (zeros outside the own shard).  Each rank keeps its own
        VidToMe pool, i.e. the multi-GPU semantics are "reference with the pool reset at shard
        boundaries"; world == 1 is exactly the reference order."""
        self._rank, self._world = int(rank), int(world)

    def _my_range(self, n: int):
        r, w = getattr(self, "_rank", 0), getattr(self, "_world", 1)
        return (n * r) // w, (n * (r + 1)) // w

    def _allreduce(self, t: torch.Tensor):
        if getattr(self, "_world", 1) > 1:
            import torch.distributed as dist
            dist.all_reduce(t)

    # ---------------------------------------------------------------- path 1
    @torch.no_grad()
    def xy_pass(self, x, conds, t, concat_conds, noises):
        """Per-chunk noise prediction over this rank's frame range (generate.py:220-224)."""
        sharded = getattr(self, "_world", 1) > 1
        f0, f1 = self._my_range(len(x))
        if sharded:
            noises.zero_()
        chunks = self.get_chunks(f1 - f0)
        vidtome.prefetch_draws(self.pipe, [len(c) for c in chunks])     # one host sync for the pass instead of one per block
        for chunk in chunks:
            lo, hi = _as_range(chunk)
            lo, hi = lo + f0, hi + f0
            cc = concat_conds[lo:hi] if concat_conds is not None else None
            self.pred_noise(x[lo:hi], conds, t, cc, batch_idx=chunk, out=noises[lo:hi])
        if sharded:
            self._allreduce(noises)
        return noises

    @torch.no_grad()
    def denoise_step(self, x, conds, conds_t, concat_conds, i: int, noises, noises_t):
        """Body of the sampling loop for step index i (generate.py:216-237)."""
        timesteps = self.scheduler._timesteps_host
        t = timesteps[i]
        self.pre_iter(x, t)
        self.xy_pass(x, conds, t, concat_conds, noises)
        if self.alpha_t > 0:
            factor = self.final_factor_t ** min(i / len(timesteps), 1)
            alpha_t = self.alpha_t * factor
            noises_t, noises = self.temporal_denoise(x, conds_t, t, concat_conds, alpha_t, noises_t, noises)
        x = self.scheduler.step(noises, t, x, generator=self.rng, return_dict=False)[0]
        self.post_iter(x, t)
        return x

    @torch.no_grad()
    def ddim_sample(self, x, conds, conds_t, concat_conds=None):
        """generate.py:207-239."""
        x = x.contiguous()
        noises = torch.zeros_like(x)
        noises_t = torch.zeros_like(x)
        for i in range(len(self.scheduler._timesteps_host)):
            x = self.denoise_step(x, conds, conds_t, concat_conds, i, noises, noises_t)
        return x

    @staticmethod
    def temporal_windows(num_frames: int, win: int):
        """Window plan of generate.py:246-260 -> (start indices, overlap list)."""
        n_slices = math.ceil((num_frames - 1) / (win - 1))
        if n_slices > 1:
            total_overlap = n_slices * win - num_frames
            overlap = total_overlap // (n_slices - 1)
            last_overlap = overlap + total_overlap % (n_slices - 1)
            overlap_list = [overlap] * (n_slices - 2) + [last_overlap]
            cum = np.cumsum(overlap_list)
            sl_idxs = [0] + [int((i + 1) * win - cum[i]) for i in range(n_slices - 1)]
        else:
            sl_idxs, overlap_list = [0], [0]
        return sl_idxs, overlap_list

    @torch.no_grad()
    def yt_pass(self, x, conds_t, t, concat_conds, noises_t, scale=None):
        """yt-plane predictions over overlapping frame windows for this rank's latent-column range
        (generate.py:246-278)."""
        scale = ops.scale_inplace if scale is None else scale
        win = self.win_size_t
        sl_idxs, overlap_list = self.temporal_windows(len(x), win)
        w0, w1 = self._my_range(x.shape[-1])
        sharded = getattr(self, "_world", 1) > 1
        if sharded:
            noises_t.zero_()
        chunks = self.get_chunks(w1 - w0)
        vidtome.prefetch_draws(self.pipe, [len(c) for _ in sl_idxs for c in chunks])
        for idx, sl_i in enumerate(sl_idxs):
            for chunk in chunks:
                c0, c1 = _as_range(chunk)
                c0, c1 = c0 + w0, c1 + w0
                # 'n c h w -> w c n h' as a view
                xt = x[sl_i:sl_i + win, :, :, c0:c1].permute(3, 1, 0, 2)
                cct = concat_conds[sl_i:sl_i + win, :, :, c0:c1].permute(3, 1, 0, 2) if concat_conds is not None else None
                out = noises_t[sl_i:sl_i + win, :, :, c0:c1].permute(3, 1, 0, 2)
                self.pred_noise(xt, conds_t, t, cct, batch_idx=chunk, sl_i=sl_i, out=out)
            if sl_i > 0:
                overlap_len = overlap_list[idx - 1]
                scale(noises_t[sl_i:sl_i + overlap_len], float(np.sqrt(0.5)))
        if sharded:
            self._allreduce(noises_t)
        return noises_t

    @torch.no_grad()
    def temporal_denoise(self, x, conds_t, t, concat_conds, alpha_t, noises_t, noises):
        """generate.py:241-284: yt-plane pass, then AdaIN to the xy statistics and blend."""
        self.yt_pass(x, conds_t, t, concat_conds, noises_t)
        ops.adain_blend(noises_t, noises, alpha_t)    # generate.py:281-282 (both updated in place)
        return noises_t, noises

    @torch.no_grad()
    def pred_noise(self, x, cond, t, concat_conds=None, batch_idx=None, sl_i=None, out=None):
        """generate.py:287-352.  ``cond`` = cat([uncond, cond]) [2, L, 768]; returns the CFG-combined
        noise [F, 4, h, w] (written into ``out`` when given)."""
        if out is None:
            out = torch.empty(tuple(x.shape), device=x.device, dtype=x.dtype)
        tt = int(t.item()) if torch.is_tensor(t) else int(t)
        self.unet.predict_noise(x, concat_conds, cond, tt, self.guidance_scale, out)
        return out

    # ---------------------------------------------------------------- path 2 (postopt.py)
    def exposure_align(self):
        from .postopt import exposure_align
        return exposure_align(self)

    def unique_tensor_optimization(self):
        from .postopt import unique_tensor_optimization
        return unique_tensor_optimization(self)

    # ---------------------------------------------------------------- pipeline level (generate.py:179-190, 552-630)
    @torch.no_grad()
    def encode_prompt_inner(self, txt: str):
        """generate.py:97-115: tokenise without truncation, cut into (model_max_length - 2)-token chunks, wrap each in
        BOS/EOS, pad with EOS, run the text encoder on all chunks -> [n_chunks, 77, 768]."""
        if self.tokenizer is None or self.text_encoder is None:
            raise TclError("Generator: pipe.tokenizer / pipe.text_encoder are required to encode prompts")
        max_length = self.tokenizer.model_max_length
        chunk_length = max_length - 2
        id_start, id_end = self.tokenizer.bos_token_id, self.tokenizer.eos_token_id
        tokens = self.tokenizer(txt, truncation=False, add_special_tokens=False)["input_ids"]
        chunks = []
        for i in range(0, len(tokens), chunk_length):
            ck = [id_start] + tokens[i:i + chunk_length] + [id_end]
            chunks.append(ck[:max_length] if len(ck) >= max_length else ck + [id_end] * (max_length - len(ck)))
        token_ids = torch.tensor(chunks).to(device=self.device, dtype=torch.int64)
        return self.text_encoder(token_ids).last_hidden_state

    @torch.no_grad()
    def encode_prompt_pair(self, positive_prompt, negative_prompt):
        """generate.py:117-135: both prompts are repeated up to the longer chunk count and their chunks concatenated
        along the token axis -> (cond, uncond), each [1, 77*k, 768] (k = 2 for the shipped prompts: L = 154)."""
        c = self.encode_prompt_inner(positive_prompt)
        uc = self.encode_prompt_inner(negative_prompt)
        max_chunk = max(len(c), len(uc))
        c = torch.cat([c] * int(math.ceil(max_chunk / len(c))), dim=0)[:max_chunk]
        uc = torch.cat([uc] * int(math.ceil(max_chunk / len(uc))), dim=0)[:max_chunk]
        return c.reshape(1, -1, c.shape[-1]), uc.reshape(1, -1, uc.shape[-1])

    @torch.no_grad()
    def encode_imgs_batch(self, imgs):
        """generate_utils.py:157-172: frames [N,3,H,W] in [0,1] -> latents (posterior mean * 0.18215)."""
        if self.vae is None:
            raise TclError("Generator: pipe.vae is required (tclight_b200.vae.AutoencoderKLB200 or a diffusers VAE)")
        if hasattr(self.vae, "encode_imgs"):
            return self.vae.encode_imgs(imgs, batch_size=self.batch_size).to(self.dtype)
        return torch.cat([self.vae.encode(2 * b.to(self.vae.dtype) - 1).latent_dist.mean * 0.18215
                          for b in imgs.split(self.batch_size, dim=0)]).to(self.dtype)

    @torch.no_grad()
    def decode_latents_batch(self, latents):
        """generate_utils.py:140-155: latents -> frames in [0,1]."""
        if self.vae is None:
            raise TclError("Generator: pipe.vae is required")
        if hasattr(self.vae, "decode_latents"):
            return self.vae.decode_latents(latents, batch_size=self.batch_size)
        return torch.cat([(self.vae.decode(1 / 0.18215 * b.to(self.vae.dtype)).sample / 2 + 0.5).clamp(0, 1)
                          for b in latents.split(self.batch_size, dim=0)])

    @torch.no_grad()
    def prepare_latents(self, n_frames: int, h: int, w: int):
        """generate.py:179-190 (IC-Light branch): one noise map repeated for every frame (noise_mode "same") or
        independent maps, scaled by the scheduler's init_noise_sigma; drawn from self.rng like
        StableDiffusionPipeline.prepare_latents does."""
        g = self.rng[0] if self.rng else None
        dev = torch.device(self.device)
        if self.noise_mode == "same":
            x = torch.randn((1, 4, h, w), generator=g, device=g.device if g is not None else dev, dtype=torch.float32)
            x = x.to(dev).repeat(n_frames, 1, 1, 1)
        else:
            x = torch.randn((n_frames, 4, h, w), generator=g, device=g.device if g is not None else dev, dtype=torch.float32).to(dev)
        return (x * self.scheduler.init_noise_sigma).to(self.dtype)

    @torch.no_grad()
    def relight(self, frames, prompt_embeds, prompt_embeds_t, future_flows=None, past_flows=None, flow_alpha=0.5,
                rgb_threshold=0.01, prepared=None):
        """The device-resident core of ``Generator.__call__`` (generate.py:560-604) for one prompt, with everything
        outside SURVEY §8 (CLIP text encoding, video decoding, the optical-flow network) supplied by the caller:

            frames [N,3,H,W] in [0,1]; prompt_embeds / prompt_embeds_t = cat([uncond, cond]) [2,L,768];
            future_flows / past_flows [N,2,H,W] (needed when post_opt.apply_opt)

        VAE-encode the frames as the IC-Light condition -> multi-axis denoising -> VAE decode -> soft masks, flow ids,
        unique inverse -> exposure alignment -> unique-video-tensor optimisation.  Returns (frames_out, info).
        ``prepared`` = the dict of ``prepare_relight`` when several prompts share one clip (the reference's prepare_data
        draws the initial noise and builds masks / unq_inv once, before its prompt loop)."""
        from . import flow_utils
        from .postopt import OptDataset

        N, _, H, W = frames.shape
        self.scheduler.set_timesteps(self.n_timesteps, device=self.device)
        if self.rng is None:
            self.rng = [torch.Generator(device=self.device).manual_seed(int(self.seed))] * N
        if prepared is None:
            prepared = self.prepare_relight(frames, future_flows, past_flows, flow_alpha, rgb_threshold)
        concat_conds, init_noise = prepared["concat_conds"], prepared["init_noise"]
        clean_latent = self.ddim_sample(init_noise, prompt_embeds, prompt_embeds_t, concat_conds)
        clean_frames = self.decode_latents_batch(clean_latent)
        info = {"latent": clean_latent}
        if self.apply_opt:
            if future_flows is None or past_flows is None:
                raise TclError("relight: post_opt.apply_opt needs future_flows and past_flows")
            masks, unq_inv = prepared["masks"], prepared["unq_inv"]
            if self.data_parser is None:
                self.data_parser = type("DataParser", (), {})()
            self.data_parser.unq_inv = unq_inv
            self.dataset = OptDataset(clean_frames.float(), past_flows.float(), masks, device=self.device)
            clean_frames, info["loss_exposure"] = self.exposure_align()
            clean_frames, info["loss_unique_tensor"] = self.unique_tensor_optimization()
            info["unq_inv"], info["mask_bwds"] = unq_inv, masks
        return clean_frames, info

    @torch.no_grad()
    def prepare_relight(self, frames, future_flows=None, past_flows=None, flow_alpha=0.5, rgb_threshold=0.01, flow_model="memflow"):
        """Per-clip inputs shared by every prompt (generate.py prepare_data): condition latents, ONE draw of the initial
        noise, and - when the optimiser runs - soft masks and the unique-tensor inverse."""
        from . import flow_utils

        N = frames.shape[0]
        if self.rng is None:
            self.rng = [torch.Generator(device=self.device).manual_seed(int(self.seed))] * N
        concat_conds = self.encode_imgs_batch(frames)
        out = {"concat_conds": concat_conds, "init_noise": self.prepare_latents(N, concat_conds.shape[2], concat_conds.shape[3])}
        if self.apply_opt:
            if future_flows is None or past_flows is None:
                raise TclError("relight: post_opt.apply_opt needs future_flows and past_flows")
            out["masks"], out["unq_inv"] = flow_utils.build_unq_inv(frames.float(), future_flows.float(), past_flows.float(),
                                                                    alpha=flow_alpha, rgb_threshold=rgb_threshold, flow_model=flow_model)
        return out

    def __call__(self, latent_path, output_path, frame_ids):
        """generate.py:560-630 with the B200 path.  Needs ``pipe.data_parser`` exposing ``load_video(frame_ids=...)`` and
        ``load_flow(frame_ids, future_flow=True, past_flow=True, gts)`` like the reference's VideoDataParser, and a text
        encoder (``self.encode_prompt_pair``) — both outside SURVEY §8 and not shipped here."""
        dp = self.data_parser
        if dp is None or not hasattr(dp, "load_video"):
            raise TclError("Generator.__call__ needs pipe.data_parser (load_video / load_flow); use relight() with tensors")
        if self.tokenizer is None or self.text_encoder is None:
            raise TclError("Generator.__call__ needs pipe.tokenizer / pipe.text_encoder; use relight() with embeddings")
        frames = dp.load_video(frame_ids=frame_ids)
        frames = frames[0] if isinstance(frames, (tuple, list)) else frames
        self.rng = [torch.Generator(device=self.device).manual_seed(int(self.seed))] * len(frame_ids)
        flows = past = None
        if self.apply_opt:
            flows, past, _ = dp.load_flow(frame_ids, True, True, frames)
        results = {}
        prepared = self.prepare_relight(frames.to(self.device), flows, past, flow_alpha=getattr(dp, "alpha", 0.5),
                                        flow_model=getattr(dp, "flow_model", "memflow"))
        for name, prompt in self.prompt.items():
            c, u = self.encode_prompt_pair(prompt, self.negative_prompt)
            ct, ut = self.encode_prompt_pair(self.prompt_t, self.negative_prompt_t)
            out, info = self.relight(frames.to(self.device), torch.cat([u, c]), torch.cat([ut, ct]), flows, past,
                                     flow_alpha=getattr(dp, "alpha", 0.5), prepared=prepared)
            results[name] = (out, info)
            if output_path:
                # generate.py:611-630: one folder per prompt with the relit video, the input video and the loss curves
                import os
                from .config_utils import save_config
                from .dataparser import save_loss_curve, save_video

                opt_suffix = "_opt" if self.apply_opt else ""
                cur = os.path.join(output_path, f"lmr_{self.local_merge_ratio}_gmr_{self.global_merge_ratio}_alpha_t_{self.alpha_t}"
                                                f"{opt_suffix}_{name}")
                save_config(self.config, cur, gene=True)
                save_video(out, cur, save_frame=self.save_frame, fps=getattr(dp, "fps", 30))
                save_video(frames, cur, save_frame=False, post_fix="_gt", fps=getattr(dp, "fps", 30))
                if self.apply_opt:
                    save_loss_curve(info["loss_exposure"], cur, "loss_exposure")
                    save_loss_curve(info["loss_unique_tensor"], cur, "loss_unique_tensor")
        return results
