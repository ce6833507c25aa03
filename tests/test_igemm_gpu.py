"""GPU parity of the tcgen05 implicit GEMM against torch fp32 references (same 16-bit inputs).

Tolerance: inputs are fp16/bf16, accumulation fp32, one rounding on output => relative L2
error <= 2e-3 (fp16) / 1e-2 (bf16) against the fp32 reference, as stated in SURVEY.md §8(d).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.float16: 2e-3, torch.bfloat16: 1e-2}


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def pack_conv_w(w):  # [Co, Ci, 3, 3] -> [Co, 9*Ci] with k = (ky*3+kx)*Ci + ci
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,K,N", [(1000, 320, 320), (4096, 640, 1280), (77, 768, 640), (300, 64, 4), (129, 1280, 960)])
def test_linear(cuda, dtype, M, K, N):
    from tclight_b200 import ops

    torch.manual_seed(0)
    x = torch.randn(M, K, device=cuda).to(dtype)
    w = (torch.randn(N, K, device=cuda) / K ** 0.5).to(dtype)
    b = torch.randn(N, device=cuda)
    r = torch.randn(M, N, device=cuda).to(dtype)
    y = ops.linear(x, w, bias=b, residual=r)
    ref = x.float() @ w.float().t() + b + r.float()
    assert rel_l2(y, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16])
@pytest.mark.parametrize("n,h,w,ci,co", [(2, 23, 40, 64, 320), (8, 12, 20, 128, 160), (1, 90, 160, 64, 64), (3, 32, 32, 320, 320)])
def test_conv3x3(cuda, dtype, n, h, w, ci, co):
    from tclight_b200 import ops

    torch.manual_seed(1)
    x = torch.randn(n, h, w, ci, device=cuda).to(dtype)
    wt = (torch.randn(co, ci, 3, 3, device=cuda) / (9 * ci) ** 0.5).to(dtype)
    b = torch.randn(co, device=cuda)
    y = ops.igemm([(x, 9, 1)], pack_conv_w(wt), (n, h, w), bias=b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < TOL[dtype]


def test_conv3x3_stride2(cuda):
    from tclight_b200 import ops

    torch.manual_seed(2)
    dtype = torch.float16
    for (n, h, w, ci, co) in [(2, 23, 40, 64, 64), (1, 90, 160, 64, 128), (2, 45, 80, 128, 64)]:
        x = torch.randn(n, h, w, ci, device=cuda).to(dtype)
        wt = (torch.randn(co, ci, 3, 3, device=cuda) / (9 * ci) ** 0.5).to(dtype)
        b = torch.randn(co, device=cuda)
        oh, ow = (h + 1) // 2, (w + 1) // 2
        y = ops.igemm([(x, 9, 2)], pack_conv_w(wt), (n, oh, ow), bias=b)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1, stride=2).permute(0, 2, 3, 1)
        assert ref.shape == y.shape
        assert rel_l2(y, ref) < TOL[dtype]


def test_resnet_tail_fused_shortcut(cuda):
    """conv2(h) + conv_shortcut(cat[x1, x2]) + bias in ONE launch (three K segments)."""
    from tclight_b200 import ops

    torch.manual_seed(3)
    dtype = torch.float16
    n, h, w = 2, 23, 40
    c1, c2, co = 128, 64, 320
    hcur = torch.randn(n, h, w, co, device=cuda).to(dtype)
    x1 = torch.randn(n, h, w, c1, device=cuda).to(dtype)
    x2 = torch.randn(n, h, w, c2, device=cuda).to(dtype)
    w2 = (torch.randn(co, co, 3, 3, device=cuda) / (9 * co) ** 0.5).to(dtype)
    ws = (torch.randn(co, c1 + c2, 1, 1, device=cuda) / (c1 + c2) ** 0.5).to(dtype)
    b = torch.randn(co, device=cuda)
    wcat = torch.cat([pack_conv_w(w2), ws.reshape(co, c1 + c2)], dim=1).contiguous()
    y = ops.igemm([(hcur, 9, 1), (x1, 1, 1), (x2, 1, 1)], wcat, (n, h, w), bias=b)
    xin = torch.cat([x1, x2], dim=-1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(hcur.float().permute(0, 3, 1, 2), w2.float(), None, padding=1) + F.conv2d(xin, ws.float())
    ref = (ref + b[None, :, None, None]).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < TOL[dtype]


def test_geglu(cuda):
    from tclight_b200 import ops, _lib as L
    from tclight_b200.weights import interleave_geglu

    torch.manual_seed(4)
    dtype = torch.float16
    M, C = 777, 320
    x = torch.randn(M, C, device=cuda).to(dtype)
    w = (torch.randn(8 * C, C, device=cuda) / C ** 0.5).to(dtype)
    b = torch.randn(8 * C, device=cuda) * 0.1
    wi, bi = interleave_geglu(w, b)
    y = ops.linear(x, wi, bias=bi, mode=L.TCL_EPI_GEGLU)
    proj = (x.float() @ w.float().t() + b).to(dtype).float()
    val, gate = proj.chunk(2, dim=-1)
    ref = val * F.gelu(gate).to(dtype).float()
    assert y.shape == (M, 4 * C)
    assert rel_l2(y, ref) < TOL[dtype]


@pytest.mark.parametrize("C,heads", [(320, 8), (640, 8), (1280, 8)])
def test_qkv_head_split(cuda, C, heads):
    from tclight_b200 import ops, _lib as L

    torch.manual_seed(5)
    dtype = torch.float16
    B, T = 2, 333
    d = C // heads
    d_pad = {40: 64, 80: 128, 160: 192}[d]
    Tp = (T + 7) // 8 * 8
    x = torch.randn(B * T, C, device=cuda).to(dtype)
    w = (torch.randn(3 * C, C, device=cuda) / C ** 0.5).to(dtype)
    q = torch.zeros(B, heads, Tp, d_pad, device=cuda, dtype=dtype)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, heads, d_pad, Tp, device=cuda, dtype=dtype)
    ops.igemm([(x.view(1, 1, B * T, C), 1, 1)], w, (1, 1, B * T), mode=L.TCL_EPI_HEADS,
              heads=dict(sec=[(q, 0), (k, 0), (vt, 1)], heads=heads, d=d, d_pad=d_pad, tok_per_batch=T, tok_pitch=Tp))
    ref = (x.float() @ w.float().t()).view(B, T, 3, heads, d)
    assert rel_l2(q[:, :, :T, :d], ref[:, :, 0].permute(0, 2, 1, 3)) < TOL[dtype]
    assert rel_l2(k[:, :, :T, :d], ref[:, :, 1].permute(0, 2, 1, 3)) < TOL[dtype]
    assert rel_l2(vt[:, :, :d, :T], ref[:, :, 2].permute(0, 2, 3, 1)) < TOL[dtype]
    assert q[:, :, T:, :].abs().max().item() == 0 and q[..., d:].abs().max().item() == 0
