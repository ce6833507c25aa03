"""CLI entry point, mirror of the reference's ``run.py``:

    python -m tclight_b200.run --config configs/my.yaml [-i video.mp4] [-p "prompt"] [--multi_axis]
    python -m tclight_b200.run --synthetic [--frames 8 --height 256 --width 256]     # random weights, synthetic clip

With real checkpoints (diffusers + IC-Light offsets on disk) it follows run.py:9-32: load_config -> seed_everything ->
init_iclight -> Generator -> generator(latents_path, output_path, frame_ids).  ``--synthetic`` runs the same device
path (VAE encode -> multi-axis denoising -> VAE decode -> masks / flow ids -> exposure alignment -> UVT optimisation) on
seeded random weights and a synthetic translating clip, and prints timings: a smoke run of the whole pipeline."""
from __future__ import annotations

import argparse
import random
import sys
import time

import numpy as np
import torch


def seed_everything(seed: int) -> None:
    """utils/VidToMe/utils.py:70-74."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def synthetic_clip(n, h, w, device, seed=0):
    """Smooth texture translating by (3, 2) px/frame with matching forward / backward flows."""
    g = torch.Generator().manual_seed(seed)
    big = torch.nn.functional.interpolate(torch.rand(1, 3, h // 8 + 8, w // 8 + 8, generator=g), size=(h + 2 * n, w + 3 * n),
                                          mode="bilinear", align_corners=False)[0]
    frames = torch.stack([big[:, 2 * (n - 1 - f):2 * (n - 1 - f) + h, 3 * (n - 1 - f):3 * (n - 1 - f) + w] for f in range(n)])
    fwd = torch.empty(n, 2, h, w)
    fwd[:, 0], fwd[:, 1] = 3.0, 2.0
    return frames.to(device), fwd.to(device), (-fwd).to(device)


def run_synthetic(a) -> int:
    from .config_utils import default_config
    from .generate import Generator
    from .model_utils import init_synthetic

    if not torch.cuda.is_available():
        print("tclight_b200.run: a CUDA device is required (there is no CPU path)", file=sys.stderr)
        return 2
    cfg = default_config(n_timesteps=a.steps, alpha_t=0.01)
    cfg.float_precision = "bf16"
    cfg.post_opt.epochs_exposure, cfg.post_opt.epochs = a.opt_epochs, a.opt_epochs
    # the reference's lr schedule spans epochs * N // batch iterations (generate.py:385): keep it >= 1 on short clips
    cfg.post_opt.batch_size = max(1, min(int(cfg.post_opt.batch_size), a.frames))
    seed_everything(cfg.seed)
    small = a.small
    pipe, scheduler, cfg.model_key = init_synthetic(
        "cuda", "bf16", unet_channels=(64, 128, 256, 256) if small else (320, 640, 1280, 1280),
        vae_channels=(64, 64, 128, 128) if small else (128, 256, 512, 512))
    gen = Generator(pipe, scheduler, cfg)
    frames, fwd, bwd = synthetic_clip(a.frames, a.height, a.width, "cuda")
    g = torch.Generator().manual_seed(1)
    conds = torch.randn(2, 154, 768, generator=g).cuda().bfloat16()
    conds_t = torch.randn(2, 77, 768, generator=g).cuda().bfloat16()
    torch.cuda.synchronize()
    t0 = time.time()
    out, info = gen.relight(frames, conds, conds_t, fwd, bwd)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"[tclight_b200] {a.frames} frames @ {a.height}x{a.width}, {a.steps} multi-axis steps + stage 1/2 "
          f"({len(info.get('loss_exposure', []))}/{len(info.get('loss_unique_tensor', []))} iterations): {dt:.2f} s; "
          f"output {tuple(out.shape)}, finite={bool(torch.isfinite(out).all())}")
    return 0


def main(argv=None) -> int:
    p = argparse.ArgumentParser(add_help=False)
    p.add_argument("--synthetic", action="store_true")
    p.add_argument("--small", action="store_true", help="tiny UNet / VAE widths (fast smoke run)")
    p.add_argument("--frames", type=int, default=8)
    p.add_argument("--height", type=int, default=256)
    p.add_argument("--width", type=int, default=256)
    p.add_argument("--steps", type=int, default=4)
    p.add_argument("--opt_epochs", type=int, default=2)
    a, rest = p.parse_known_args(argv)
    if a.synthetic:
        return run_synthetic(a)
    # ---- reference flow (run.py:9-32) ----
    from .config_utils import load_config
    from .dataparser import VideoDataParser, get_frame_ids
    from .generate import Generator
    from .model_utils import init_iclight

    config = load_config(argv=rest)
    seed_everything(config.seed)
    if config.sd_version != "iclight":
        raise SystemExit("tclight_b200.run: only sd_version 'iclight' (the TC-Light path) is wired here; "
                         "use tclight_b200.invert.Inverter for the DDIM-inversion path")
    pipe, scheduler, config.model_key = init_iclight(config.device)
    pipe.data_parser = VideoDataParser(config.data, config.device)
    generator = Generator(pipe, scheduler, config)
    frame_ids = get_frame_ids(config.generation.frame_range, pipe.data_parser.n_frames, config.generation.frame_ids)
    generator(config.generation.latents_path, config.generation.output_path, frame_ids=frame_ids)
    return 0


if __name__ == "__main__":
    sys.exit(main())
