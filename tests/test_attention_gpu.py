"""GPU parity of the tcgen05 attention kernel vs torch fp32 softmax attention on the same
16-bit inputs.  Tolerance: P is rounded to 16 bit before P·V (as in flash attention), output
rounded once => rel-L2 <= 3e-3 (fp16) / 1.5e-2 (bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


import contextlib


@contextlib.contextmanager
def _variant(variant):
    """variant -1 = the product library; the tuning variants live in libtclight_tuning.so (include/tclight_tuning.h), which
    is swapped in for ops.attention while the block runs."""
    from tclight_b200 import _lib, ops

    if variant == -1:
        yield
        return
    t = _lib.load_tuning_lib()
    if t is None:
        pytest.skip("libtclight_tuning.so not built (make -C tclight_b200/csrc tuning)")
    old_lib, old_var = ops.lib, t.tcl_debug_attention_variant(variant)
    ops.lib = t
    try:
        yield
    finally:
        ops.lib = old_lib
        t.tcl_debug_attention_variant(old_var)
TOL = {torch.float16: 3e-3, torch.bfloat16: 1.5e-2}


def _mk(B, H, T, d, d_pad, dtype, dev, scale=1.0):
    Tp = (T + 7) // 8 * 8
    x = (torch.randn(B, H, T, d, device=dev) * scale).to(dtype)
    pad = torch.zeros(B, H, Tp, d_pad, device=dev, dtype=dtype)
    pad[:, :, :T, :d] = x
    return x, pad


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,Tq,Tk,d,div", [
    (2, 8, 300, 300, 40, 1),     # self-attention, ragged tiles
    (2, 8, 1024, 1024, 40, 1),
    (1, 8, 257, 640, 80, 1),
    (2, 8, 130, 130, 160, 1),
    (4, 8, 200, 154, 40, 2),     # cross-attention: 2 frames per CFG half share text K/V
    (4, 8, 200, 77, 80, 2),
    (2, 4, 64, 77, 16, 1),       # tiny-UNet head dim
])
def test_attention(cuda, dtype, B, H, Tq, Tk, d, div):
    from tclight_b200 import ops

    torch.manual_seed(0)
    d_pad = ops.head_pad(d)
    q, qp = _mk(B, H, Tq, d, d_pad, dtype, cuda, 1.5)
    k, kp = _mk(B // div, H, Tk, d, d_pad, dtype, cuda, 1.5)
    v, vp = _mk(B // div, H, Tk, d, d_pad, dtype, cuda)
    vt = vp.transpose(2, 3).contiguous()
    out = ops.attention(qp, kp, vt, Tq, Tk, d, kv_batch_div=div)
    kk = k.float().repeat_interleave(div, dim=0)
    vv = v.float().repeat_interleave(div, dim=0)
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), kk) / d ** 0.5
    ref = torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), vv).permute(0, 2, 1, 3).reshape(B, Tq, H * d)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    assert err < TOL[dtype], err


@pytest.mark.parametrize("variant", [0, 5, 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_variants(cuda, dtype, variant):
    """Every tuning variant of the kernel (scalar / packed-pair arithmetic, FMA-pipe exp2 share, stale reference) is held
    to the same tolerance as the shipped one."""
    from tclight_b200 import ops

    with _variant(variant):
        for (B, H, Tq, Tk, d, div) in [(2, 8, 1024, 1024, 40, 1), (2, 8, 300, 300, 40, 1), (1, 8, 257, 640, 80, 1),
                                       (2, 8, 130, 130, 160, 1), (4, 8, 200, 154, 40, 2)]:
            torch.manual_seed(0)
            d_pad = ops.head_pad(d)
            q, qp = _mk(B, H, Tq, d, d_pad, dtype, cuda, 1.5)
            k, kp = _mk(B // div, H, Tk, d, d_pad, dtype, cuda, 1.5)
            v, vp = _mk(B // div, H, Tk, d, d_pad, dtype, cuda)
            out = ops.attention(qp, kp, vp.transpose(2, 3).contiguous(), Tq, Tk, d, kv_batch_div=div)
            kk = k.float().repeat_interleave(div, dim=0)
            vv = v.float().repeat_interleave(div, dim=0)
            s = torch.einsum("bhqd,bhkd->bhqk", q.float(), kk) / d ** 0.5
            ref = torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), vv).permute(0, 2, 1, 3).reshape(B, Tq, H * d)
            err = ((out.float() - ref).norm() / ref.norm()).item()
            assert err < TOL[dtype], (variant, d, err)


@pytest.mark.parametrize("variant", [-1, 0, 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("boost", [4.0, 12.0, 40.0])
def test_attention_row_max_jumps_between_tiles(cuda, dtype, variant, boost):
    """Keys of the later KV tiles are `boost` times larger, so row maxima jump by 2^10 ... 2^100 between tiles: exercises
    the lazy rescale, the deferred rescale of the stale-reference variants and their overflow slow path (the scores are
    re-read from TMEM and the tile is redone against the new reference)."""
    from tclight_b200 import ops

    with _variant(variant):
        B, H, Tq, Tk, d = 1, 8, 300, 700, 40
        torch.manual_seed(3)
        d_pad = ops.head_pad(d)
        q, qp = _mk(B, H, Tq, d, d_pad, dtype, cuda, 1.5)
        k, kp = _mk(B, H, Tk, d, d_pad, dtype, cuda, 1.0)
        v, vp = _mk(B, H, Tk, d, d_pad, dtype, cuda)
        scale = torch.ones(Tk, device=cuda)
        scale[128:] = boost
        scale[384:] = boost * 2
        scale[640:] = boost * 4
        k = (k.float() * scale[None, None, :, None]).to(dtype)
        kp = torch.zeros_like(kp)
        kp[..., :Tk, :d] = k
        out = ops.attention(qp, kp, vp.transpose(2, 3).contiguous(), Tq, Tk, d)
        s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) / d ** 0.5
        ref = torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v.float()).permute(0, 2, 1, 3).reshape(B, Tq, H * d)
        assert bool(torch.isfinite(out.float()).all())
        err = ((out.float() - ref).norm() / ref.norm()).item()
        assert err < TOL[dtype], (variant, boost, err)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("T", [4800, 4700, 5000])
def test_attention_kv_split_tail(cuda, dtype, T):
    """Launches whose CTA count leaves a small partial last wave (here 8 heads x 19-20 Q tile groups = 152-160 work items on
    148 SMs) run that tail as KV-split CTAs + a merge kernel: the result must meet the same tolerance, including a ragged last
    key tile (T = 4700, 5000) and rows beyond tq in the last Q tile."""
    from tclight_b200 import ops

    B, H, d = 1, 8, 40
    torch.manual_seed(1)
    d_pad = ops.head_pad(d)
    q, qp = _mk(B, H, T, d, d_pad, dtype, cuda, 1.5)
    k, kp = _mk(B, H, T, d, d_pad, dtype, cuda, 1.5)
    v, vp = _mk(B, H, T, d, d_pad, dtype, cuda)
    out = ops.attention(qp, kp, vp.transpose(2, 3).contiguous(), T, T, d)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, T, H * d)
    assert bool(torch.isfinite(out.float()).all())
    err = ((out.float() - ref).norm() / ref.norm()).item()
    # the rows of the tail items specifically (the last work items in launch order = the last Q tile groups of the last head)
    tail_rows = slice(T - 256, T)
    err_tail = ((out.float()[:, tail_rows, -d:] - ref[:, tail_rows, -d:]).norm() / ref[:, tail_rows, -d:].norm()).item()
    assert err < TOL[dtype] and err_tail < TOL[dtype], (T, err, err_tail)
