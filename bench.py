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
    p.add_argument("--no-ref-on-b200", action="store_true")
    p.add_argument("--no-aux", action="store_true")
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
# The reference's PyTorch op sequence (oracle/unet_ref.py + the oracle VidToMe patch), used as the BASELINE that is
# timed, never as the product: on the host cores (cpu_baseline / --impl reference) and on the B200 itself
# (reference_on_b200: cuDNN / cuBLAS / SDPA with materialised VidToMe scores).
# ----------------------------------------------------------------------------------------------
class SdpaFlops:
    """torch's FlopCounterMode does not see the CPU scaled_dot_product_attention: count 4*B*H*Tq*Tk*d here."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig, self.flops = F, F.scaled_dot_product_attention, 0.0

        def counted(q, k, v, *a, **kw):
            self.flops += 4.0 * q.shape[0] * q.shape[1] * q.shape[2] * k.shape[2] * q.shape[3]
            return self.orig(q, k, v, *a, **kw)

        F.scaled_dot_product_attention = counted
        return self

    def __exit__(self, *exc):
        self.F.scaled_dot_product_attention = self.orig


def oracle_chunks(device, dtype, h, w, plane, seed=0):
    """Returns run(k): the k-th chunk-forward of one denoising step through the oracle UNet with the oracle ToMe patch
    (xy: 4 frames of h x w, L=154; yt: 4 latent columns as 64 x h 'images', L=77), CFG pair, guidance 2.0.  Chunk 0 has no
    global-token pool, later chunks merge against it (the steady state of a pass)."""
    import torch
    from oracle import pipeline_ref as P
    from oracle.unet_ref import apply_oracle_patch, make_unet, reset_oracle_pool

    unet = make_unet(seed=0).to(device=device, dtype=dtype)
    apply_oracle_patch(unet, 0.6, True, 0.5, global_rand=0.5)
    g = torch.Generator().manual_seed(seed)
    ih, iw = (h, w) if plane == "xy" else (64, h)
    L = 154 if plane == "xy" else 77
    x = torch.randn(1, 4, ih, iw, generator=g).repeat(12, 1, 1, 1).to(device=device, dtype=dtype)
    cc = (0.18215 * (torch.randn(1, 4, ih, iw, generator=g) + 0.02 * torch.randn(12, 4, ih, iw, generator=g))).to(device=device, dtype=dtype)
    text = torch.randn(2, L, 768, generator=g).to(device=device, dtype=dtype)
    t = torch.tensor(801, device=device)

    def run(k):
        sl = slice(4 * (k % 3), 4 * (k % 3) + 4)
        with torch.no_grad():
            return P.cfg_noise(unet, x[sl], text, t, cc[sl], 2.0)

    run.reset = lambda: reset_oracle_pool(unet)
    return run


def cpu_path1_sample(h, w):
    """Two consecutive 4-frame xy chunk-forwards (without / with the global-token pool) of the oracle UNet + ToMe patch in
    fp32 on all host threads.  Returns (seconds, TFLOP executed)."""
    import torch
    from torch.utils.flop_counter import FlopCounterMode

    torch.set_num_threads(os.cpu_count() or 1)
    run = oracle_chunks("cpu", torch.float32, h, w, "xy")
    with SdpaFlops() as sd, FlopCounterMode(display=False) as fc:
        t0 = time.perf_counter()
        run(0)
        run(1)
        sec = time.perf_counter() - t0
    return sec, (fc.get_total_flops() + sd.flops) / 1e12


def cpu_stage2_sample(H, W, frames=16, iters=2):
    """Stage-2 iterations of the oracle optimiser (torch autograd, the reference's op sequence) on the host cores: a
    `frames`-frame clip at the bench resolution, batch 16.  Returns seconds per iteration (init subtracted)."""
    import torch
    from oracle import postopt_ref as O
    from tclight_b200.postopt import synthetic_workload

    torch.set_num_threads(os.cpu_count() or 1)
    edited, flows, masks, inv = synthetic_workload(frames, H, W, "cpu")
    batch = [list(range(min(16, frames)))]
    t0 = time.perf_counter()
    O.stage2_uvt(edited, flows, masks, inv, [batch])
    t1 = time.perf_counter()
    O.stage2_uvt(edited, flows, masks, inv, [batch * (1 + iters)])
    t2 = time.perf_counter()
    return max(((t2 - t1) - (t1 - t0)) / iters, 1e-9)


def step_units(frames, w_lat, win=64, chunk=4):
    import math
    n_xy = math.ceil(frames / chunk)
    n_win = max(1, math.ceil((frames - 1) / (win - 1)))
    n_yt = n_win * math.ceil(w_lat / chunk)
    return n_xy, n_yt


def step_tflop_model(frames, h, w):
    """TFLOP of one full multi-axis step by the FLOP model of SURVEY.md §8(d) (scaled by pixel count off 720x1280)."""
    n_xy, n_yt = step_units(frames, w)
    px = (h * w) / (90.0 * 160.0)
    return (n_xy * XY_CHUNK_TF + n_yt * YT_CHUNK_TF * (min(frames, 64) / 64.0)) * px


def cpu_baseline(args, step_tflop=None, with_stage2=True):
    """The reference's CPU path timed on this box's host cores, on a bounded sample: the denoising path as TFLOP/s of real
    4-frame xy chunk-forwards with the ToMe patch at HALF the latent resolution (a full-resolution chunk-forward is ~64
    TFLOP, minutes on a CPU), converted to steps/s with the step's FLOPs; stage 2 as it/s of the oracle optimiser."""
    h, w = args.height // 8, args.width // 8
    hs, ws = max(h // 2, 16), max(w // 2, 16)
    sec, tf = cpu_path1_sample(hs, ws)
    rate = tf / sec
    step_tf = step_tflop if step_tflop else step_tflop_model(args.frames, h, w)
    out = {"value": rate / step_tf, "unit": "steps/s", "cores": os.cpu_count() or 1, "kind": "port",
           "sample": (f"oracle UNet + ToMe patch fp32: two 4-frame xy chunk-forwards (no pool / pool) at latent {hs}x{ws} = {tf:.2f} TFLOP in "
                      f"{sec:.1f} s = {rate:.3f} TFLOP/s; one full step = {step_tf:.0f} TFLOP"
                      f" ({'counted from the B200 arm launches' if step_tflop else 'FLOP model of SURVEY.md 8d'})"),
           "path1_cpu_tflops": rate}
    if with_stage2:
        try:
            s_it = cpu_stage2_sample(args.height, args.width)
            out["stage2"] = {"value": 1.0 / s_it, "unit": "it/s",
                             "sample": f"oracle stage-2 optimiser (torch autograd fp32), 16 frames at {args.height}x{args.width}, batch 16, 2 iterations"}
        except Exception as ex:  # noqa: BLE001
            out["stage2"] = {"value": None, "error": str(ex)[:200]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h, w = args.height // 8, args.width // 8
    hs, ws = max(h // 2, 16), max(w // 2, 16)
    step_tf = step_tflop_model(args.frames, h, w)
    rates, secs = [], []
    for i in range(min(args.warmup, 1) + max(1, min(args.steps, 4))):      # each "step" = one bounded sample; capped to stay in minutes
        sec, tf = cpu_path1_sample(hs, ws)
        if i >= min(args.warmup, 1):
            rates.append(tf / sec)
            secs.append(sec)
    rate = sum(rates) / len(rates)
    sps = rate / step_tf
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "denoising_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": (f"oracle UNet + ToMe patch fp32: two 4-frame xy chunk-forwards (no pool / pool) at latent {hs}x{ws}, "
                                    f"{len(rates)} sample(s) of {sum(secs) / len(secs):.1f} s = {rate:.3f} TFLOP/s; one full step = {step_tf:.0f} TFLOP "
                                    "(FLOP model of SURVEY.md 8d)")},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        s_it = cpu_stage2_sample(args.height, args.width)
        line["stage2"] = {"metric": "stage2_iters_per_sec", "value": 1.0 / s_it, "unit": "it/s",
                          "sample": f"oracle stage-2 optimiser (torch autograd fp32), 16 frames at {args.height}x{args.width}, batch 16, 2 iterations"}
    except Exception as ex:  # noqa: BLE001
        line["stage2"] = {"value": None, "error": str(ex)[:200]}
    print(json.dumps(line), flush=True)


def workload_config(args):
    """Identical for both arms (the driver compares the two lines' configs)."""
    h, w = args.height // 8, args.width // 8
    n_xy, n_yt = step_units(args.frames, w)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"workload": f"{args.frames}f@{args.height}x{args.width} multi-axis denoising + VidToMe (chunk 4, mix-4, 0.6/0.5), L=154/77, guidance 2.0",
            "unit_def": "step = one full-video multi-axis denoising step", "xy_chunk_forwards_per_step": n_xy,
            "yt_chunk_forwards_per_step": n_yt,
            "parallelism": f"frame/column shards x{world}, all-reduce of noises" if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (activations of one chunk-forward exceed 126 MB)", "weights": "seeded random, SD-1.5 shapes (860M)"}


def tracked_traffic():
    """dram bytes per launch of the roofline kernels, read from the tracked ncu summary (profiles/traffic.json, written by
    tools/ncu_traffic.py from the committed ncu csv) instead of a literal."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def reference_on_b200(dev, dtype, h, w, frames, n_xy, n_yt, mine_xy_ms, mine_yt_ms):
    """The reference's PyTorch path on this B200: steady-state xy and yt chunk-forwards of the oracle
    UNet + ToMe patch (cuDNN convs, cuBLAS linears, SDPA attention, materialised matching scores) in the bench dtype, CUDA
    events, extrapolated to a step by the call counts."""
    import torch

    out = {}
    for plane, n_calls in (("xy", n_xy), ("yt", n_yt)):
        try:
            run = oracle_chunks(dev, dtype, h, w, plane)
            run(0); run(1)                      # warm-up: cuDNN autotune, pool built
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            run(2); run(0)
            e.record()
            torch.cuda.synchronize()
            out[plane + "_chunk_forward_ms"] = s.elapsed_time(e) / 2
            del run
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            out[plane + "_error"] = str(ex)[:200]
    if "xy_chunk_forward_ms" in out and "yt_chunk_forward_ms" in out:
        step_ms = n_xy * out["xy_chunk_forward_ms"] + n_yt * out["yt_chunk_forward_ms"]
        out.update({"steps_per_sec": 1e3 / step_ms, "ms_per_step": step_ms,
                    "this_repo_xy_chunk_forward_ms": mine_xy_ms, "this_repo_yt_chunk_forward_ms": mine_yt_ms,
                    "note": "oracle/unet_ref.py + oracle ToMe patch on CUDA (library kernels), steady-state chunk-forwards x call counts; "
                            "sampler / AdaIN / scheduler time not included on the reference side"})
    return out


def reference_stage2_on_b200(dev, H, W, frames=16, iters=3):
    """torch-autograd stage-2 iterations (oracle/postopt_ref.py, the reference's op sequence) on this B200."""
    import torch
    from oracle import postopt_ref as O
    from tclight_b200.postopt import synthetic_workload

    edited, flows, masks, inv = synthetic_workload(frames, H, W, dev)
    batch = [list(range(min(16, frames)))]

    def timed(k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.stage2_uvt(edited, flows, masks, inv, [batch * k])
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    timed(1)
    a, b = timed(1), timed(1 + iters)
    return {"value": iters / max(b - a, 1e-9), "unit": "it/s",
            "sample": f"torch autograd fp32 on CUDA, {frames} frames at {H}x{W}, batch 16 (Adam over that clip's U rows only)"}


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
    e2e_steps = max(2, min(args.steps, 5))                 # the same step with host copies; 5 steps bound the run time
    ms_e2e = timed(e2e_step, e2e_steps)
    clocks = clk.stop() if rank == 0 else None
    finite = bool(torch.isfinite(state["x"].float()).all().item())

    # the two plane passes on their own (xy and yt chunk-forwards/s separately)
    t_mid = timesteps[len(timesteps) // 2]
    ms_xy = timed(lambda: (gen.pre_iter(state["x"], t_mid), gen.xy_pass(state["x"], conds, t_mid, state["cc"], state["noises"]),
                           gen.post_iter(state["x"], t_mid)), 1)
    ms_yt = timed(lambda: (gen.pre_iter(state["x"], t_mid), gen.yt_pass(state["x"], conds_t, t_mid, state["cc"], state["noises_t"]),
                           gen.post_iter(state["x"], t_mid)), 1)

    pk, pk_kind = peaks()
    traffic = tracked_traffic()
    n_xy, n_yt = step_units(N, w)
    sps = args.steps / (ms * 1e-3)
    sps_e2e = e2e_steps / (ms_e2e * 1e-3)
    # dominant kernel: ds-1 self-attention: algorithmic FLOPs / event time
    a = prof.get("attention_d64", {"flops": 0.0, "ms": 0.0, "launches": 0})
    roof = None
    if a["ms"] > 0:
        ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
        tr = traffic.get("attention_ds1", {})
        roof = {"bound": "tensor", "kernel": "attn_kernel (ds-1 self/cross attention, head dim 40)", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "peak_kind": f"{pk_kind} sustained bf16", "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                "traffic_unit": "bytes/launch (ncu dram__bytes_read+write, largest launch shape: merged xy self-attention, T=47520)",
                "launches": a["launches"], "avg_launch_ms": a["ms"] / max(1, a["launches"]),
                "share_of_step": a["ms"] / ms}
    total_fl = sum(v["flops"] for v in prof.values())
    cfgd = workload_config(args)
    line = {
        "metric": "denoising_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": cfgd,
        "chunk_forwards_per_sec": (n_xy + n_yt) * sps,
        "xy_pass": {"ms": ms_xy, "chunk_forwards": n_xy, "chunk_forwards_per_sec": n_xy / (ms_xy * 1e-3)},
        "yt_pass": {"ms": ms_yt, "chunk_forwards": n_yt, "chunk_forwards_per_sec": n_yt / (ms_yt * 1e-3)},
        # per-rank path FLOPs: with shards the first chunk of every shard has no pool yet, so N ranks do slightly LESS attention
        # work than one (the driver's scaling efficiency should be read next to these)
        "path_tflop_per_step_all_ranks": None,
        "path_tflops": total_fl * world / (ms * 1e-3) / 1e12 if total_fl else None,
        "path_frac_of_sustained_bf16": (total_fl / (ms * 1e-3) / 1e12) / pk["bf16_tflops_sustained"] if total_fl else None,
        "kernel_breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        "roofline": roof,
        "e2e": {"value": sps_e2e, "unit": "steps/s", "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size() + cc_host.numel() * cc_host.element_size()),
                "d2h_bytes_per_step": int(x_host.numel() * x_host.element_size()), "steps": e2e_steps},
        "gpu_launches": int(launches.item()), "clocks": clocks, "finite": finite,
    }
    fl_t = torch.tensor([total_fl / max(1, args.steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(fl_t)
    line["path_tflop_per_step_all_ranks"] = fl_t.item() / 1e12
    # free the denoising state before the other legs
    del state, unet, gen, pipe
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_ref_on_b200:
        try:
            line["reference_on_b200"] = reference_on_b200(dev, adt, h, w, N, n_xy, n_yt, ms_xy / n_xy, ms_yt / n_yt)
        except Exception as ex:  # noqa: BLE001
            line["reference_on_b200"] = {"error": str(ex)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args, step_tflop=line["path_tflop_per_step_all_ranks"] or None)
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
    if rank == 0 and not args.no_aux:
        try:
            line["aux"] = aux_legs(dev, adt, H, W)
        except Exception as ex:  # noqa: BLE001
            line["aux"] = {"error": str(ex)[:200]}
        torch.cuda.empty_cache()
    if args.stage2_iters > 0:
        from tclight_b200 import postopt
        line["stage2"], line["stage1"] = postopt.bench_postopt(dev, N, H, W, iters=args.stage2_iters, rank=rank, world=world,
                                                               traffic=traffic.get("stage2_iteration", {}))
        if rank == 0 and world == 1 and not args.no_ref_on_b200:
            try:
                line["stage2"]["reference_on_b200"] = reference_stage2_on_b200(dev, H, W)
                line["stage2"]["convergence_vs_oracle"] = stage2_convergence_vs_oracle(dev)
            except Exception as ex:  # noqa: BLE001
                line["stage2"]["reference_on_b200"] = {"error": str(ex)[:200]}
    # the second half of the headline metric at the top level too, so a per-N curve of the lines shows it
    if isinstance(line.get("stage2"), dict):
        line["stage2_iters_per_sec"] = line["stage2"].get("value")
    if isinstance(line.get("stage1"), dict) and line["stage1"]:
        line["stage1_iters_per_sec"] = line["stage1"].get("value")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def aux_legs(dev, adt, H, W):
    """The callers either side of the denoising path at the bench resolution (BASELINE config 5 names them):
    VAE encode / decode (frames/s, batch 4) and one DDIM-inversion step of `Inverter` (un-patched UNet, no CFG, batch 8,
    invert.py:151-173).  Per-GPU rates on seeded random SD-1.5-shaped weights; frames shard trivially across ranks."""
    import types

    import torch
    from tclight_b200.invert import Inverter
    from tclight_b200.scheduler import DDIMSchedulerB200
    from tclight_b200.unet import UNetB200
    from tclight_b200.vae import AutoencoderKLB200
    from tclight_b200.weights import random_state_dict, random_vae_state_dict

    def ms_of(fn, warm=1, k=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / k

    out = {}
    h, w = H // 8, W // 8
    vae = AutoencoderKLB200(random_vae_state_dict(seed=0), device=dev, dtype=adt)
    lat = (0.18215 * torch.randn(4, 4, h, w, device=dev)).to(adt)
    img = torch.rand(4, 3, H, W, device=dev)
    out["vae_decode_frames_per_sec"] = 4e3 / ms_of(lambda: vae.decode_latents(lat))
    out["vae_encode_frames_per_sec"] = 4e3 / ms_of(lambda: vae.encode_imgs(img))
    del vae, lat, img
    torch.cuda.empty_cache()

    class D(dict):
        __getattr__ = dict.__getitem__

    fp = "fp16" if adt == torch.float16 else "bf16"
    cfg = types.SimpleNamespace(device="cuda", sd_version="1.5", model_key=None, float_precision=fp, height=H, width=W, work_dir=".",
                                inversion=D(float_precision=fp, control="none", control_scale=1.0, save_steps=50, steps=50, prompt="",
                                            recon=False, save_intermediate=False, use_blip=False, batch_size=8, force=True, n_frames=None))
    unet = UNetB200(random_state_dict(seed=0, in_channels=4), device=dev, dtype=adt)
    inv = Inverter(types.SimpleNamespace(unet=unet), DDIMSchedulerB200(), cfg)
    x = torch.randn(8, 4, h, w, device=dev).to(inv.dtype if hasattr(inv, "dtype") else adt)
    conds = torch.randn(8, 77, 768, device=dev).to(adt)
    ts = list(reversed(inv.scheduler.timesteps))

    def one():
        eps = inv._all_noise(x, conds, ts[3])
        inv.pred_next_x(x, eps, ts[3], 3, inversion=True)

    ms = ms_of(one)
    out["inversion_frame_steps_per_sec"] = 8e3 / ms
    out["inversion_note"] = ("Inverter: one DDIM-inversion step of 8 frames (un-patched SD-1.5 UNet forward, L=77, no CFG, + tcl_ddim_next); "
                             "a clip needs frames x 50 such frame-steps")
    return out


def stage2_convergence_vs_oracle(dev):
    """The full reference budget (70 epochs) of stage 2 on a down-scaled clip through this repo's optimiser and through the
    oracle (torch autograd), same batches: loss-curve end points side by side."""
    import types

    import torch
    from oracle import postopt_ref as O
    from tclight_b200.postopt import OptDataset, unique_tensor_optimization

    n, hh, ww, bo, epochs = 16, 176, 192, 8, 70
    edited, flows, masks, inv = O.synthetic_clip(n=n, h=hh, w=ww, seed=1, device=dev)
    ds = OptDataset(edited.clone(), flows, masks, device=dev)
    gen = types.SimpleNamespace(dataset=ds, data_parser=types.SimpleNamespace(unq_inv=inv), lambda_dssim=0.2, lambda_flow=0.8,
                                lambda_tv=0.05, epochs_exposure=0, epochs=epochs, opt_batch_size=bo, feature_lr=0.05,
                                exposure_lr_init=0.01, exposure_lr_final=0.001, exposure_lr_delay_steps=0, exposure_lr_delay_mult=0.0)
    torch.manual_seed(7)
    _, got = unique_tensor_optimization(gen)
    torch.manual_seed(7)
    _, _, want = O.stage2_uvt(edited, flows, masks, inv, O.draw_batches(n, bo, epochs), feature_lr=0.05)
    return {"clip": f"{n} frames {hh}x{ww}, batch {bo}, {epochs} epochs = {len(got)} iterations",
            "loss_first": [got[0], want[0]], "loss_last": [got[-1], want[-1]], "order": "[this repo, oracle]",
            "max_abs_diff_over_curve": max(abs(a - b) for a, b in zip(got, want))}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
