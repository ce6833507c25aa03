"""2-GPU parity check of the data-parallel optimiser (launched by tests/test_dp_postopt_2gpu.py, or by hand on a 2-GPU box:
`python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_postopt_check_2gpu.py`); every rank runs the DP
path (stage 2: UVT rows sharded over peer memory; stage 1: all-reduced exposure gradient), rank 0 also runs the single-GPU path
and compares losses / images.  N = 7 frames over 2 ranks (N % world != 0), and the ranks' CPU RNGs are deliberately driven apart
before each stage, as the sharded denoising passes do (get_chunks consumes a shard-length-dependent amount of RNG)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import postopt_ref as O
from tclight_b200 import postopt as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def mk(ds, inv, world_, rank_):
    g = types.SimpleNamespace(dataset=ds, data_parser=types.SimpleNamespace(unq_inv=inv), lambda_dssim=0.2, lambda_flow=0.8, lambda_tv=0.05,
                              epochs_exposure=2, epochs=2, opt_batch_size=4, feature_lr=0.05, exposure_lr_init=0.01, exposure_lr_final=0.001,
                              exposure_lr_delay_steps=0, exposure_lr_delay_mult=0.0, _world=world_, _rank=rank_)
    return g

edited, flows, masks, inv = O.synthetic_clip(n=7, h=176, w=192, seed=1, device=dev)
for stage in (2, 1):
    torch.manual_seed(5)
    if rank > 0:
        torch.randperm(2 + rank)            # diverged CPU RNG on the other ranks
    ds = P.OptDataset(edited.clone(), flows, masks, device=dev)
    fn = P.unique_tensor_optimization if stage == 2 else P.exposure_align
    img_dp, loss_dp = fn(mk(ds, inv, world, rank))
    if rank == 0:
        torch.manual_seed(5)
        ds1 = P.OptDataset(edited.clone(), flows, masks, device=dev)
        img_1, loss_1 = fn(mk(ds1, inv, 1, 0))
        dl = max(abs(a - b) for a, b in zip(loss_dp, loss_1) if a == a and b == b)
        d = (img_dp - img_1).abs()
        print(f"stage {stage}: DP(x{world}) vs single GPU: {len(loss_dp)} iterations, max loss diff {dl:.2e}, image diff mean {d.mean().item():.2e} max {d.max().item():.2e}")
        assert dl < 1e-5 and d.mean().item() < 5e-4
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("dp postopt ok")
