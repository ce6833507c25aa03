"""GPU parity of the VAE path (tclight_b200.vae.AutoencoderKLB200: tcl_igemm convs incl. the (0,1,0,1)-padded
stride-2 downsampler, tcl_groupnorm, the three-launch single-head attention with tcl_softmax_rows, and the staging
kernels) against the fp32 oracle restatement of diffusers' AutoencoderKL (oracle/vae_ref.py) on the same weights.
Tolerance: 16-bit activations through ~30 layers => rel-L2 <= 1e-2 (fp16) / 4e-2 (bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
SMALL = dict(block_out_channels=(64, 64, 128, 128))


def _rel(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 4e-2)])
@pytest.mark.parametrize("hw", [(64, 96), (72, 88)])       # latent 8x12 (T=96 -> padded keys) and 9x11 (T=99)
def test_vae_encode_decode_vs_oracle(cuda, dtype, tol, hw):
    from oracle import vae_ref as V
    from tclight_b200.vae import AutoencoderKLB200

    ref = V.make_vae(seed=0, **SMALL)
    mine = AutoencoderKLB200(ref.state_dict(), device=cuda, dtype=dtype, **SMALL)
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(3, 3, *hw, generator=g)
    want_lat = V.encode_imgs(ref, imgs)
    got_lat = mine.encode_imgs(imgs.to(cuda))
    assert got_lat.shape == want_lat.shape and got_lat.dtype == dtype
    assert _rel(got_lat, want_lat) < tol
    want_img = V.decode_latents(ref, want_lat)
    got_img = mine.decode_latents(want_lat.to(cuda))
    assert got_img.shape == want_img.shape
    assert _rel(got_img, want_img) < tol
    assert float(got_img.min()) >= 0.0 and float(got_img.max()) <= 1.0
    # diffusers-style operators
    with torch.no_grad():
        m_ref = ref.encode(2 * imgs - 1).latent_dist.mean
        s_ref = ref.decode(want_lat / 0.18215).sample
    assert _rel(mine.encode((2 * imgs - 1).to(cuda).to(dtype)).latent_dist.mean, m_ref) < tol
    assert _rel(mine.decode((want_lat / 0.18215).to(cuda).to(dtype)).sample, s_ref) < tol


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_softmax_rows(cuda, dtype):
    from tclight_b200._lib import check, dtype_code, lib, stream_ptr

    torch.manual_seed(0)
    for rows, cols, pitch in [(5, 99, 128), (33, 256, 256), (7, 1003, 1024)]:
        x = (torch.randn(rows, pitch, device=cuda) * 3).to(dtype)
        want = torch.zeros(rows, pitch)
        want[:, :cols] = x[:, :cols].float().softmax(-1).cpu()
        check(lib.tcl_softmax_rows(dtype_code(dtype), x.data_ptr(), rows, cols, pitch, stream_ptr()), "softmax")
        assert (x.float().cpu() - want).abs().max() < (2e-3 if dtype == torch.float16 else 8e-3)
        assert float(x[:, cols:].abs().max()) == 0.0 if cols < pitch else True


def test_asymmetric_pad_stride2_conv(cuda):
    """tcl_igemm no_lead_pad == F.conv2d(F.pad(x, (0,1,0,1)), w, stride=2)."""
    import torch.nn.functional as F
    from tclight_b200 import ops
    from tclight_b200.weights import pack_conv3x3

    torch.manual_seed(0)
    for h, w in [(16, 24), (18, 10)]:
        x = torch.randn(2, 64, h, w)
        wt = torch.randn(64, 64, 3, 3) * 0.05
        want = F.conv2d(F.pad(x, (0, 1, 0, 1)), wt, stride=2)
        xn = x.permute(0, 2, 3, 1).contiguous().to(cuda).half()
        got = ops.igemm([(xn, 9, 2, True)], pack_conv3x3(wt).to(cuda).half(), (2, want.shape[2], want.shape[3]))
        assert _rel(got.permute(0, 3, 1, 2), want) < 3e-3
