"""Stage-2 (Unique-Video-Tensor) iteration benchmark at the BASELINE shape, single GPU or under torchrun.
python tools/bench_stage2.py [--frames 300 --height 720 --width 1280 --iters 40]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

p = argparse.ArgumentParser()
p.add_argument("--frames", type=int, default=300)
p.add_argument("--height", type=int, default=720)
p.add_argument("--width", type=int, default=1280)
p.add_argument("--iters", type=int, default=40)
a = p.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from tclight_b200.postopt import bench_stage2
r = bench_stage2(dev, a.frames, a.height, a.width, iters=a.iters, rank=rank, world=world)
if rank == 0:
    print(json.dumps(r))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
