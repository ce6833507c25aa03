"""One steady-state xy chunk-forward and one steady-state yt chunk-forward at the C3 shapes (720x1280: 4 frames x 90x160;
4 latent columns x 64 frames x 90 rows), each bracketed by marker launches, for an ncu launch list:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/chunk.csv python tools/prof_chunk.py
  python tools/prof_chunk.py --summarise gpurun_out/chunk.csv      # per-kernel shares, weighted 75 xy + 200 yt per step

Markers are tcl scale_kernel launches on a 1-element tensor (no such launch occurs inside a chunk-forward)."""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run():
    import numpy as np
    import torch
    from tclight_b200 import ops
    from tclight_b200.config_utils import default_config
    from tclight_b200.generate import Generator
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200
    from tclight_b200.unet import UNetB200
    from tclight_b200.weights import random_state_dict

    dev = torch.device("cuda")
    adt = torch.bfloat16
    unet = UNetB200(random_state_dict(seed=0), device=dev, dtype=adt)
    cfg = default_config(alpha_t=0.01)
    cfg.float_precision = "bf16"
    pipe = type("Pipe", (), {})()
    pipe.unet = unet
    gen = Generator(pipe, DPMSolverMultistepSchedulerB200(), cfg)
    torch.manual_seed(0); np.random.seed(0)
    N, h, w = 72, 90, 160
    x = torch.randn(N, 4, h, w, device=dev).to(adt)
    cc = (0.18215 * torch.randn(N, 4, h, w, device=dev)).to(adt)
    conds = torch.randn(2, 154, 768, device=dev).to(adt)
    conds_t = torch.randn(2, 77, 768, device=dev).to(adt)
    out = torch.zeros_like(x)
    mark = torch.ones(1, device=dev, dtype=adt)
    t = int(gen.scheduler._timesteps_host[0])
    # xy: chunk 0 seeds the pool, chunk 1 is the warm-up of the steady state, chunk 2 is measured
    for k in range(3):
        if k == 2:
            ops.scale_inplace(mark, 1.0)
        gen.pred_noise(x[4 * k:4 * k + 4], conds, t, cc[4 * k:4 * k + 4], out=out[4 * k:4 * k + 4])
    ops.scale_inplace(mark, 1.0)
    gen.post_iter(x, t)
    for k in range(3):
        xt = x[:64, :, :, 4 * k:4 * k + 4].permute(3, 1, 0, 2)
        cct = cc[:64, :, :, 4 * k:4 * k + 4].permute(3, 1, 0, 2)
        ot = out[:64, :, :, 4 * k:4 * k + 4].permute(3, 1, 0, 2)
        if k == 2:
            ops.scale_inplace(mark, 1.0)
        gen.pred_noise(xt, conds_t, t, cct, out=ot)
    ops.scale_inplace(mark, 1.0)
    torch.cuda.synchronize()


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    ix = {h: i for i, h in enumerate(rows[hdr])}
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) < len(ix):
            continue
        v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
        v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        grid = r[ix["Grid Size"]] if "Grid Size" in ix else ""
        seq.append((r[ix["Kernel Name"]].split("(")[0], v, grid))
    marks = [i for i, (n, v, g) in enumerate(seq) if "scale_kernel" in n and g.replace(" ", "") in ("(1,1,1)", "1,1,1", "1")]
    assert len(marks) >= 4, f"markers found: {len(marks)}"
    seg = {"xy": seq[marks[0] + 1:marks[1]], "yt": seq[marks[2] + 1:marks[3]]}
    weight = {"xy": 75, "yt": 200}
    total = collections.defaultdict(float)
    for k, s in seg.items():
        tot = sum(v for _, v, _ in s)
        print(f"# {k} chunk-forward: {len(s)} launches, {tot / 1e3:.2f} ms (cold-cache, serialised under ncu)")
        for n, v, _ in s:
            total[n] += v * weight[k]
    gt = sum(total.values())
    print(f"# step = 75 xy + 200 yt chunk-forwards: {gt / 1e6:.2f} s of kernel time under ncu")
    for n, v in sorted(total.items(), key=lambda kv: -kv[1])[:30]:
        print(f"{n[:84]:84s} {v / 1e3:10.1f} ms {100 * v / gt:5.1f}%")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run()
