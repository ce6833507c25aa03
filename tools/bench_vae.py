"""VAE decode / encode throughput at 720x1280 with SD-1.5-sized random weights (frames/s, batch 2 like the reference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tclight_b200.vae import AutoencoderKLB200
from tclight_b200.weights import random_vae_state_dict

dev = torch.device("cuda")
sd = random_vae_state_dict(seed=0)
vae = AutoencoderKLB200(sd, device=dev, dtype=torch.bfloat16)
lat = (0.18215 * torch.randn(4, 4, 90, 160, device=dev)).to(torch.bfloat16)
img = torch.rand(4, 3, 720, 1280, device=dev)
for name, fn in [("decode_latents", lambda: vae.decode_latents(lat)), ("encode_imgs", lambda: vae.encode_imgs(img))]:
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = fn(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{name}: 4 frames @720x1280 in {ms:.1f} ms = {4e3/ms:.1f} frames/s; finite={bool(torch.isfinite(out.float()).all())}, shape={tuple(out.shape)}")
