"""GPU parity of the HBM-bound helper kernels vs torch fp32 references.
Tolerances: 16-bit output rounding => rel-L2 <= 2e-3 (fp16); index/copy kernels are bit-exact."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("n,h,w,c1,c2,silu,eps", [(2, 23, 40, 320, 0, True, 1e-5), (8, 12, 20, 1280, 1280, True, 1e-5),
                                                  (2, 45, 80, 640, 320, True, 1e-5), (3, 16, 16, 64, 0, False, 1e-6),
                                                  (2, 8, 8, 256, 128, True, 1e-5)])
def test_groupnorm(cuda, n, h, w, c1, c2, silu, eps):
    from tclight_b200 import ops

    torch.manual_seed(0)
    dt = torch.float16
    x1 = (torch.randn(n, h, w, c1, device=cuda) * 2 + 0.5).to(dt)
    x2 = (torch.randn(n, h, w, c2, device=cuda) * 0.5 - 1).to(dt) if c2 else None
    C = c1 + c2
    g = torch.randn(C, device=cuda)
    b = torch.randn(C, device=cuda)
    y = ops.groupnorm(x1, g, b, 32, eps, silu, x2=x2)
    xin = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.group_norm(xin.float().permute(0, 3, 1, 2), 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < 2e-3


@pytest.mark.parametrize("rows,C", [(1000, 320), (77, 640), (513, 1280), (64, 64)])
def test_layernorm(cuda, rows, C):
    from tclight_b200 import ops

    torch.manual_seed(1)
    x = (torch.randn(rows, C, device=cuda) * 3 + 1).half()
    g = torch.randn(C, device=cuda)
    b = torch.randn(C, device=cuda)
    y = ops.layernorm(x, g, b, 1e-5)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    assert rel_l2(y, ref) < 2e-3


@pytest.mark.parametrize("h,w,oh,ow", [(12, 20, 23, 40), (23, 40, 45, 80), (45, 80, 90, 160), (8, 8, 16, 16), (8, 12, 16, 23)])
def test_upsample_nearest(cuda, h, w, oh, ow):
    from tclight_b200 import ops

    x = torch.randn(2, h, w, 64, device=cuda).half()
    y = ops.upsample_nearest(x, oh, ow)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), size=(oh, ow), mode="nearest").permute(0, 2, 3, 1).half()
    assert torch.equal(y, ref)


@pytest.mark.parametrize("ldt", [torch.float16, torch.float32])
def test_stage_and_cfg(cuda, ldt):
    from tclight_b200 import ops
    from einops import rearrange

    torch.manual_seed(2)
    N, h, w = 10, 12, 20
    x = torch.randn(N, 4, h, w, device=cuda).to(ldt)
    cond = torch.randn(N, 4, h, w, device=cuda).to(ldt)
    # xy chunk
    st = ops.stage_latent(x[2:5], cond[2:5], torch.float16)
    ref = torch.cat([x[2:5], cond[2:5]], 1).permute(0, 2, 3, 1).half()
    assert torch.equal(st[:3, ..., :8], ref) and torch.equal(st[3:, ..., :8], ref) and st[..., 8:].abs().max() == 0
    # yt view (generate.py:267): 'n c h w -> w c n h'
    xv = rearrange(x[1:9, :, :, 4:8], "n c h w -> w c n h")
    cv = rearrange(cond[1:9, :, :, 4:8], "n c h w -> w c n h")
    st = ops.stage_latent(xv, cv, torch.float16)
    ref = torch.cat([xv, cv], 1).permute(0, 2, 3, 1).half()
    assert torch.equal(st[:4, ..., :8], ref)
    # cfg store into the strided view
    eps = torch.randn(8, 8, h, 8, device=cuda).half()
    tgt = torch.zeros_like(x)
    tv = rearrange(tgt[1:9, :, :, 4:8], "n c h w -> w c n h")
    ops.cfg_store(eps, 2.0, tv)
    e = eps[..., :4].permute(0, 3, 1, 2).to(ldt)
    u, c = e[:4], e[4:]
    want = u + 2.0 * (c - u)
    got = rearrange(tgt[1:9, :, :, 4:8], "n c h w -> w c n h")
    assert torch.allclose(got.float(), want.float(), atol=2e-3, rtol=2e-3)
    assert tgt[0].abs().max() == 0 and tgt[:, :, :, :4].abs().max() == 0
