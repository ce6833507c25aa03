"""Full-size (BASELINE.json configs[2]: 720x1280, 4-frame chunk) checks of the CUDA path through
size-independent properties and sub-sampled comparisons, where running the whole oracle would take minutes:
  * VidToMe at the ds-1 shapes: merge/unmerge index algebra (round trip, replace semantics, pool update);
  * attention at T = 47 520: rows of softmax sum to one (V = 1 => O = 1) and a random subset of query rows
    against fp32 softmax attention;
  * implicit GEMM conv at [8, 90, 160, 320] against torch conv2d (fp32) on one image;
  * one stage-2 iteration at 720x1280 against the oracle's autograd loss and gradient.
"""
import copy
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _Mod:
    pass


def test_vidtome_c3_shapes_and_roundtrip(cuda):
    from tclight_b200 import ops
    from tclight_b200.vidtome import patch

    torch.manual_seed(0)
    Fn, n, Cc = 4, 14400, 320
    args = dict(max_downsample=2, generator=None, seed=123, batch_size=2, align_batch=True, merge_global=True,
                global_merge_ratio=0.5, local_merge_ratio=0.6, global_rand=0.5, target_stride=4)
    info = dict(size=(90, 160), args=copy.deepcopy(args))
    mod = _Mod()
    mod.generator = torch.Generator(device=cuda).manual_seed(1)
    shapes = []
    for chunk in range(2):
        base = torch.randn(2, 1, n, Cc, device=cuda)
        x = (base + 0.3 * torch.randn(2, Fn, n, Cc, device=cuda)).reshape(2 * Fn, n, Cc).half()
        m, u, merged, plan = patch.compute_merge_plan(mod, x, info)
        shapes.append(tuple(merged.shape))
        back = u(merged)                                           # [8, n, C]
        xj = x.reshape(2, Fn * n, Cc)
        bj = back.reshape(2, Fn * n, Cc)
        same = (bj == xj).all(dim=-1)                              # tokens that survived keep their exact value
        # every restored token is an actual token of the merged sequence (replace mode copies, never blends)
        um = plan.total_unmerge_map.long()
        assert torch.equal(bj, merged[:, um])
        frac = same.float().mean().item()
        assert 0.2 < frac < 0.9, frac
        assert mod.global_tokens.shape == (2, 31680, Cc)
    # first chunk [2, 31680, 320], steady state [2, 47520, 320]
    assert shapes == [(2, 31680, 320), (2, 47520, 320)]


def test_attention_c3_rowsum_and_subset(cuda):
    from tclight_b200 import ops

    torch.manual_seed(1)
    B, H, T, d = 2, 8, 47520, 40
    dp, Tp = ops.head_pad(d), T
    dt = torch.float16
    q = torch.zeros(B, H, Tp, dp, device=cuda, dtype=dt)
    k = torch.zeros(B, H, Tp, dp, device=cuda, dtype=dt)
    q[..., :d] = (torch.randn(B, H, T, d, device=cuda) * 1.5).to(dt)
    k[..., :d] = (torch.randn(B, H, T, d, device=cuda) * 1.5).to(dt)
    vt = torch.zeros(B, H, dp, Tp, device=cuda, dtype=dt)
    vt[:, :, :d] = 1.0
    out = ops.attention(q, k, vt, T, T, d)
    assert (out.float() - 1.0).abs().max().item() < 2e-3            # softmax rows sum to one
    v = (torch.randn(B, H, T, d, device=cuda)).to(dt)
    vt[:, :, :d] = v.transpose(2, 3)
    out = ops.attention(q, k, vt, T, T, d).view(B, T, H, d)
    # 64 random rows + 32 rows of the last Q tile groups: at this shape (2 976 work items on 148 SMs) the last 16 items of the
    # last (batch, head) run as the KV-split tail (9 slices + merge kernel)
    rows = torch.cat([torch.randint(0, T, (64,), device=cuda), torch.randint(T - 16 * 256, T, (32,), device=cuda)])
    qs = q[:, :, rows, :d].float()
    s = torch.einsum("bhqd,bhkd->bhqk", qs, k[..., :d].float()) / d ** 0.5
    ref = torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v.float()).permute(0, 2, 1, 3)
    err = ((out[:, rows].float() - ref).norm() / ref.norm()).item()
    assert err < 3e-3, err
    err_tail = ((out[1:, rows[64:], H - 1].float() - ref[1:, 64:, H - 1]).norm() / ref[1:, 64:, H - 1].norm()).item()
    assert err_tail < 3e-3, err_tail


def test_conv_c3_shape(cuda):
    from tclight_b200 import ops
    from tclight_b200.weights import pack_conv3x3

    torch.manual_seed(2)
    x = torch.randn(8, 90, 160, 320, device=cuda).half()
    w = (torch.randn(320, 320, 3, 3, device=cuda) / (9 * 320) ** 0.5).half()
    b = torch.randn(320, device=cuda)
    y = ops.igemm([(x, 9, 1)], pack_conv3x3(w), (8, 90, 160), bias=b)
    for img in (0, 7):
        ref = F.conv2d(x[img:img + 1].float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
        err = ((y[img:img + 1].float() - ref).norm() / ref.norm()).item()
        assert err < 2e-3, err


def test_stage2_iteration_720p(cuda):
    from oracle import postopt_ref as O
    from tclight_b200 import postopt as P
    from tclight_b200._lib import lib, check, stream_ptr

    n, h, w = 4, 720, 1280
    edited, flows, masks, inv = O.synthetic_clip(n=n, h=h, w=w, seed=5, device="cpu")
    edited, flows, masks, inv = (t.to(cuda) for t in (edited, flows, masks, inv))
    ds = P.OptDataset(edited, flows, masks, device=cuda)
    idx = [2, 0, 3]
    size = int(inv.max().item()) + 1
    mean_rgb = O.scatter_mean(edited.permute(0, 2, 3, 1).reshape(-1, 3), inv, size)
    fdc0 = ((mean_rgb - 0.5) / O.SH_C0 + 0.1 * torch.randn(size, 3, device=cuda)).contiguous()
    fdc = fdc0.clone().requires_grad_(True)
    idx_t = torch.tensor(idx, device=cuda)
    both = torch.cat([idx_t, (idx_t - 1).clamp(min=0)])
    rgb = torch.index_select(fdc * O.SH_C0 + 0.5, 0, inv.reshape(n, h, w)[both].reshape(-1)).clamp(0, 1)
    out = rgb.reshape(len(both), h, w, 3).permute(0, 3, 1, 2)
    img, pre = out[:3], out[3:]
    flow = O._flow_term(img, pre, flows[idx_t], masks[idx_t], idx_t)
    photo = (1 - O.ms_ssim_relaxed(img, edited[idx_t])) * 0.2
    loss = 0.2 * photo + 0.8 * flow + O.tv_loss(img, 0.05)
    loss.backward()
    ctx = P._Context(ds, 0.2, 0.8, 0.05, 3)
    ids = inv.to(torch.int32).contiguous()
    p = fdc0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    g = torch.zeros((p.shape[0], 4), device=cuda)        # UVT gradient rows are {dR, dG, dB, pad}
    lo = torch.zeros(3, device=cuda)
    arr = (C.c_int * 3)(*idx)
    check(lib.tcl_uvt_iteration(C.byref(ctx.c), arr, 3, ids.data_ptr(), size, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                0.0, 0.9, 0.999, 1e-15, 1, lo.data_ptr(), stream_ptr()), "uvt")
    grad = m / 0.1
    err = (grad - fdc.grad).abs()
    outl = (err > 1e-2 * fdc.grad.abs().max()).float().mean().item()
    good = err <= 1e-2 * fdc.grad.abs().max()
    rel = ((grad - fdc.grad)[good].norm() / fdc.grad[good].norm()).item()
    print(f"720p stage-2 iteration: loss {lo[0].item():.7f} vs {loss.item():.7f}; gradient inlier rel-L2 {rel:.2e}, outliers {outl:.1e}")
    assert abs(lo[0].item() - loss.item()) < 1e-5 and rel < 5e-3 and outl < 1e-3
