#!/usr/bin/env python
"""bench.py — headline benchmark of the two TC-Light hot paths on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 300 frames at
720x1280 (latent 90x160), multi-axis denoising with VidToMe (chunk 4, mix-4, ratios 0.6/0.5),
synthetic latents/conditions, seeded random SD-1.5-shaped weights, bf16 activations.
A "step" = ONE full-video multi-axis denoising step: 75 xy chunk-forwards + 5 windows x 40 column
chunk-forwards through the UNet, AdaIN/blend, DPM-Solver++ update, pool reset.
`value` = steps/s with inputs resident in HBM; `e2e` = the same step through Generator with the
step's latents/conditions copied from pinned host memory and the new latent read back, inside
the timed region.  Stage-2 iterations/s (the second half of the metric) are reported in
`stage2` when that path is built.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# FLOP model of SURVEY.md §8(d) / BASELINE.md §3 (per chunk-forward, 720x1280, steady state)
XY_CHUNK_TF, YT_CHUNK_TF, PLAIN_IMAGE_TF = 63.6, 14.8, 3.94


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--frames", type=int, default=300)
    p.add_argument("--height", type=int, default=720)
    p.add_argument("--width", type=int, default=1280)
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--stage2-iters", type=int, default=20)
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port timed on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample(h, w, repeats=1, warm=0):
    """One xy chunk-forward of a single frame (CFG pair = 2 images, text L=154) through the oracle
    UNet (torch fp32, all host threads).  Returns seconds per sample."""
    import torch
    from oracle.unet_ref import make_unet

    torch.set_num_threads(os.cpu_count() or 1)
    unet = make_unet(seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, h, w, generator=g)
    cc = torch.randn(1, 4, h, w, generator=g) * 0.18215
    text = torch.randn(2, 154, 768, generator=g)
    t = torch.tensor(801)
    times = []
    with torch.no_grad():
        for i in range(warm + repeats):
            t0 = time.perf_counter()
            unet(torch.cat([x, x]), t, encoder_hidden_states=text, cross_attention_kwargs={"concat_conds": cc})
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return times


def step_units(frames, w_lat, win=64, chunk=4):
    import math
    n_xy = math.ceil(frames / chunk)
    n_win = max(1, math.ceil((frames - 1) / (win - 1)))
    n_yt = n_win * math.ceil(w_lat / chunk)
    return n_xy, n_yt


def cpu_extrapolate(sec_per_sample, frames, h, w):
    """steps/s of a full multi-axis step, extrapolated from the single-frame sample by the FLOP
    model (scaled by pixel count when not at 720x1280)."""
    n_xy, n_yt = step_units(frames, w)
    px = (h * w) / (90.0 * 160.0)
    step_tf = (n_xy * XY_CHUNK_TF + n_yt * YT_CHUNK_TF * (min(frames, 64) / 64.0)) * px
    sample_tf = 2 * PLAIN_IMAGE_TF * px
    return 1.0 / (sec_per_sample * step_tf / sample_tf), step_tf / sample_tf


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h, w = args.height // 8, args.width // 8
    times = cpu_sample(h, w, repeats=max(1, args.steps), warm=min(args.warmup, 1))
    sec = sum(times) / len(times)
    sps, ratio = cpu_extrapolate(sec, args.frames, h, w)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "denoising_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.frames}f@{args.height}x{args.width} multi-axis denoising + VidToMe (chunk 4, mix-4, 0.6/0.5), L=154/77, guidance 2.0",
                   "unit_def": "step = one full-video multi-axis denoising step",
                   "note": "reference is Python/PyTorch: timed = oracle port of its CPU path (oracle/unet_ref.py), bounded sample extrapolated by the FLOP model"},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"1-frame xy chunk-forward (2 images, L=154) = {sec:.2f} s; extrapolated x{ratio:.0f} by the FLOP model"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tclight_b200 import _lib as L
    from tclight_b200 import ops
    from tclight_b200.config_utils import default_config
    from tclight_b200.generate import Generator
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200
    from tclight_b200.unet import UNetB200
    from tclight_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    adt = torch.bfloat16 if args.dtype == "bf16" else torch.float16

    N, H, W = args.frames, args.height, args.width
    h, w = H // 8, W // 8
    sd = random_state_dict(seed=0)
    unet = UNetB200(sd, device=dev, dtype=adt)
    del sd
    cfg = default_config(alpha_t=0.01)
    cfg.float_precision = "fp16" if adt == torch.float16 else "bf16"
    pipe = type("Pipe", (), {})()
    pipe.unet = unet
    gen = Generator(pipe, DPMSolverMultistepSchedulerB200(), cfg)
    gen.set_shard(rank, world)
    ldt = gen.dtype

    # synthetic inputs: same-noise latents, temporally smooth condition latents
    torch.manual_seed(12345)
    torch.cuda.manual_seed(12345)
    np.random.seed(12345)
    g = torch.Generator().manual_seed(12345)
    x_host = torch.randn(1, 4, h, w, generator=g).repeat(N, 1, 1, 1).to(ldt).pin_memory()
    base = torch.randn(1, 4, h, w, generator=g)
    drift = torch.cumsum(torch.randn(N, 4, h, w, generator=g) * 0.02, dim=0)
    cc_host = (0.18215 * (base + drift)).to(ldt).pin_memory()
    conds = torch.randn(2, 154, 768, generator=g).to(adt).to(dev)
    conds_t = torch.randn(2, 77, 768, generator=g).to(adt).to(dev)
    gen.rng = [torch.Generator(device=dev).manual_seed(12345)] * N
    timesteps = gen.scheduler._timesteps_host

    state = {"x": x_host.to(dev), "cc": cc_host.to(dev), "i": 0,
             "noises": torch.zeros(N, 4, h, w, device=dev, dtype=ldt), "noises_t": torch.zeros(N, 4, h, w, device=dev, dtype=ldt)}

    def one_step(x, cc):
        i = state["i"] % len(timesteps)
        if i == 0:
            gen.scheduler.set_timesteps(cfg.generation.n_timesteps, device=dev)
        out = gen.denoise_step(x, conds, conds_t, cc, i, state["noises"], state["noises_t"])
        state["i"] += 1
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def dev_step():
        state["x"] = one_step(state["x"], state["cc"])

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        cd = cc_host.to(dev, non_blocking=True)
        out = one_step(xd, cd)
        x_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        dev_step()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    L.lib.tcl_launch_count_reset()
    ops.profile_start()
    ms = timed(dev_step, args.steps)
    prof = ops.profile_stop()
    launches = torch.tensor([L.lib.tcl_launch_count()], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)
    ms_e2e = timed(e2e_step, args.steps)
    clocks = clk.stop() if rank == 0 else None
    finite = bool(torch.isfinite(state["x"].float()).all().item())

    pk, pk_kind = peaks()
    n_xy, n_yt = step_units(N, w)
    sps = args.steps / (ms * 1e-3)
    sps_e2e = args.steps / (ms_e2e * 1e-3)
    # dominant kernel: ds-1 self-attention (attn_kernel<2,64,..>): algorithmic FLOPs / event time
    a = prof.get("attention_d64", {"flops": 0.0, "ms": 0.0, "launches": 0})
    roof = None
    if a["ms"] > 0:
        ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "attn_kernel<NQ=2,DPAD=64,TS,POLY=0x03> (ds-1 self/cross attention)", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "peak_kind": f"{pk_kind} sustained bf16", "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"],
                # dram__bytes_read+write of ONE launch of this kernel at its largest shape (ds-1 merged xy self-attention,
                # B*H=16, T=47 520): profiles/r01_attention_v2_ncu_summary.md; algorithmic Q,K,V^T,O bytes of that launch: 353 MB
                "traffic": 347.0e6, "traffic_unit": "bytes/launch (ncu --set full, largest launch shape)",
                "launches": a["launches"], "avg_launch_ms": a["ms"] / max(1, a["launches"]),
                "share_of_step": a["ms"] / ms}
    total_fl = sum(v["flops"] for v in prof.values())
    line = {
        "metric": "denoising_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{N}f@{H}x{W} multi-axis denoising + VidToMe (chunk 4, mix-4, 0.6/0.5), L=154/77, guidance 2.0",
                   "unit_def": "step = one full-video multi-axis denoising step", "xy_chunk_forwards_per_step": n_xy,
                   "yt_chunk_forwards_per_step": n_yt, "parallelism": f"frame/column shards x{world}, all-reduce of noises" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (activations of one chunk-forward exceed 126 MB)", "weights": "seeded random, SD-1.5 shapes (860M)"},
        "chunk_forwards_per_sec": (n_xy + n_yt) * sps,
        "path_tflops": total_fl * world / (ms * 1e-3) / 1e12 if total_fl else None,
        "path_frac_of_sustained_bf16": (total_fl / (ms * 1e-3) / 1e12) / pk["bf16_tflops_sustained"] if total_fl else None,
        "kernel_breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        "roofline": roof,
        "e2e": {"value": sps_e2e, "unit": "steps/s", "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size() + cc_host.numel() * cc_host.element_size()),
                "d2h_bytes_per_step": int(x_host.numel() * x_host.element_size())},
        "gpu_launches": int(launches.item()), "clocks": clocks, "finite": finite,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            times = cpu_sample(h, w, repeats=1, warm=0)
            v, ratio = cpu_extrapolate(times[0], N, h, w)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"oracle UNet fp32, 1-frame xy chunk-forward (2 images) = {times[0]:.2f} s, extrapolated x{ratio:.0f} by FLOP model"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
    try:
        from tclight_b200 import postopt
        if hasattr(postopt, "bench_stage2") and args.stage2_iters > 0:
            line["stage2"] = postopt.bench_stage2(dev, N if world == 1 else N, H, W, iters=args.stage2_iters, rank=rank, world=world)
    except ImportError:
        pass
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
