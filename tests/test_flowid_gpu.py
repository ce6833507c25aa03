"""GPU parity of the stage-2 producer kernels (csrc/flowid.cu through tclight_b200.flow_utils) against
oracle/flowid_ref.py and the goldens recorded from the reference (tests/golden/flowid_producer.pt).
Integer outputs (flow ids, unique inverse) are bit-exact; fp32 warps within 2e-5, soft masks within 1e-4 (fma
contraction in the bicubic taps / norms is amplified by beta = 100 inside the sigmoid; expf vs the CPU's vectorised exp)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("seed", [0, 1])
def test_producer_vs_reference_golden(cuda, seed):
    from oracle import flowid_ref as R
    from tclight_b200 import flow_utils as F

    want = torch.load(os.path.join(GOLD, "flowid_producer.pt"))[seed]
    frames, fwd, bwd = R.synthetic_scene(n=6, h=40, w=56, seed=seed)
    fr, fw, bw = frames.to(cuda), fwd.to(cuda), bwd.to(cuda)
    masks = F.get_soft_mask_bwds(fr * 2 - 1, fw, bw, alpha=0.5)
    assert masks.shape == want["masks"].shape
    assert (masks.cpu() - want["masks"]).abs().max() < 1e-4
    warp = F.warp_flow(fr, bw)
    assert (warp.cpu() - want["warp"]).abs().max() < 2e-5
    # ids from the reference's mask (identical input => identical integers)
    ids, cnt = F.get_flowid(fr, fw, want["masks"].to(cuda), rgb_threshold=0.05, return_count=True)
    assert ids.dtype == torch.int32 and torch.equal(ids.cpu(), want["ids"])
    assert int(cnt) == int(want["ids"].max()) + 1
    inv = F.voxelization(ids.view(-1, 1), fr.permute(0, 2, 3, 1).reshape(-1, 3), None, None)
    assert inv.dtype == torch.int64 and torch.equal(inv.cpu(), want["inv"].reshape(-1))


@pytest.mark.parametrize("n,h,w,seed,thr", [(5, 36, 44, 2, 0.01), (3, 64, 48, 3, 0.2), (1, 24, 40, 4, 0.05), (9, 51, 67, 5, 0.05)])
def test_producer_vs_oracle(cuda, n, h, w, seed, thr):
    from oracle import flowid_ref as R
    from tclight_b200 import flow_utils as F

    frames, fwd, bwd = R.synthetic_scene(n=n, h=h, w=w, seed=seed)
    m_ref = R.soft_mask_bwds(frames, fwd, bwd, alpha=0.3, diff_threshold=0.05)
    fr, fw, bw = frames.to(cuda), fwd.to(cuda), bwd.to(cuda)
    m = F.get_soft_mask_bwds(fr, fw, bw, alpha=0.3, diff_threshold=0.05)
    assert (m.cpu() - m_ref).abs().max() < 1e-4
    ids = F.get_flowid(fr, fw, m_ref.to(cuda), rgb_threshold=thr)
    assert torch.equal(ids.cpu(), R.flow_ids(frames, fwd, m_ref, rgb_threshold=thr))
    # arbitrary (sparse, repeated, unsorted) ids through the general unique-inverse path
    g = torch.Generator().manual_seed(seed)
    sparse = torch.randint(0, 5000, (n * h * w,), generator=g, dtype=torch.int32) * 7 + 3
    assert torch.equal(F.voxelization(sparse.to(cuda).view(-1, 1)).cpu(), R.unique_inverse(sparse))


def test_build_unq_inv_feeds_stage2(cuda):
    """The producer's output is what unique_tensor_optimization consumes: dense int64 ids, frame 0 = arange."""
    from oracle import flowid_ref as R
    from tclight_b200 import flow_utils as F

    frames, fwd, bwd = R.synthetic_scene(n=6, h=40, w=56, seed=7)
    masks, inv = F.build_unq_inv(frames.to(cuda), fwd.to(cuda), bwd.to(cuda), alpha=0.5, rgb_threshold=0.05)
    m_ref = R.soft_mask_bwds(frames * 2 - 1, fwd, bwd, alpha=0.5)
    assert (masks.cpu() - m_ref).abs().max() < 1e-4
    P = 40 * 56
    assert torch.equal(inv[:P].cpu(), torch.arange(P))
    U = int(inv.max()) + 1
    assert torch.equal(torch.unique(inv).cpu(), torch.arange(U))
    # pixels the mask is far from deciding differently: same ids as the oracle run on the device mask
    assert torch.equal(inv.cpu().view(6, 40, 56).to(torch.int32), R.flow_ids(frames, fwd, masks.cpu(), rgb_threshold=0.05))


def test_producer_full_size_properties(cuda):
    """720x1280, 24 frames: size-independent properties (ids dense, frame 0 = arange, inherited ids point at a
    pixel of the previous frame with the same id within the rounded flow, inverse is idempotent)."""
    from tclight_b200 import flow_utils as F

    N, H, W = 24, 720, 1280
    g = torch.Generator(device=cuda).manual_seed(0)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, H // 8 + 8, W // 8 + 8, device=cuda, generator=g),
                                           size=(H + 2 * N, W + 3 * N), mode="bilinear")[0]
    frames = torch.stack([base[:, 2 * (N - 1 - f):2 * (N - 1 - f) + H, 3 * (N - 1 - f):3 * (N - 1 - f) + W] for f in range(N)])
    fwd = torch.empty(N, 2, H, W, device=cuda)
    fwd[:, 0], fwd[:, 1] = 3.0, 2.0
    bwd = -fwd
    masks, inv = F.build_unq_inv(frames, fwd, bwd, alpha=0.5, rgb_threshold=0.01)
    P = H * W
    ids = inv.view(N, H, W)
    assert torch.equal(ids[0].reshape(-1), torch.arange(P, device=cuda))
    U = int(inv.max()) + 1
    present = torch.zeros(U, dtype=torch.bool, device=cuda)
    present[inv] = True
    assert bool(present.all())                                    # dense
    # interior pixels follow the exact (3, 2) translation: id(f, y, x) == id(f-1, y-2, x-3)
    assert torch.equal(ids[1:, 8:, 8:], ids[:-1, 6:-2, 5:-3])
    assert masks[1:, :, 8:, 8:].min() > 0.99 and float(masks[0].min()) == 1.0
    assert U < 0.2 * N * P
    again = F.voxelization(inv.to(torch.int32).view(-1, 1), id_range=U)
    assert torch.equal(again, inv)                                # idempotent on dense ids
