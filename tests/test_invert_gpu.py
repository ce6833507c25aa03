"""GPU parity of the DDIM-inversion path (tclight_b200.invert.Inverter + tcl_ddim_next) against the oracle
restatement of the reference's Inverter (oracle/pipeline_ref.ddim_walk, pinned bit-exactly to invert.py on CPU)."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(steps=5, batch=4, fp="fp16"):
    inv = dict(float_precision=fp, control="none", control_scale=1.0, save_steps=steps, steps=steps, prompt="", recon=False,
               save_intermediate=False, use_blip=False, batch_size=batch, force=True, n_frames=None)

    class D(dict):
        __getattr__ = dict.__getitem__

    return types.SimpleNamespace(device="cuda", sd_version="1.5", model_key=None, inversion=D(inv), float_precision=fp,
                                 height=128, width=128, work_dir=".")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_ddim_next_matches_torch_op_sequence(cuda, dtype):
    """tcl_ddim_next == the reference's six tensor ops in the latent dtype, bit for bit (16-bit) / 1 ulp (fp32 fma)."""
    from tclight_b200.invert import Inverter
    from tclight_b200.scheduler import DDIMSchedulerB200

    I = Inverter(types.SimpleNamespace(unet=None), DDIMSchedulerB200(), _cfg())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 4, 18, 22, generator=g).to(dtype).to(cuda)
    eps = torch.randn(7, 4, 18, 22, generator=g).to(dtype).to(cuda)
    sch = I.scheduler
    for inversion in (True, False):
        ts = reversed(sch.timesteps) if inversion else sch.timesteps
        for i in (0, 2, len(ts) - 1):
            t = ts[i]
            got = I.pred_next_x(x, eps, t, i, inversion=inversion)
            a_t = sch.alphas_cumprod[t]
            if inversion:
                a_p = sch.alphas_cumprod[ts[i - 1]] if i > 0 else sch.final_alpha_cumprod
            else:
                a_p = sch.alphas_cumprod[ts[i + 1]] if i < len(ts) - 1 else sch.final_alpha_cumprod
            mu, sg, mu_p, sg_p = a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5
            if inversion:
                want = mu * ((x - sg_p * eps) / mu_p) + sg * eps
            else:
                want = mu_p * ((x - sg * eps) / mu) + sg_p * eps
            assert want.dtype == dtype
            if dtype == torch.float32:
                assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
            else:
                assert torch.equal(got, want)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 5e-2)])
def test_inverter_vs_oracle(cuda, dtype, tol):
    from oracle import make_goldens as G, pipeline_ref as P
    from oracle.scheduler_ref import DDIMRef
    from oracle.unet_ref import make_unet
    from tclight_b200.invert import Inverter
    from tclight_b200.scheduler import DDIMSchedulerB200
    from tclight_b200.unet import UNetB200

    ref_unet = make_unet(seed=0, **G.INV_UNET)
    x, conds = G.inversion_inputs()
    sch = DDIMRef()
    sch.set_timesteps(5)
    with torch.no_grad():
        want = P.ddim_walk(ref_unet, sch, x, conds, 4, True)
        want0 = P.ddim_walk(ref_unet, sch, want, conds, 4, False)
    unet = UNetB200(ref_unet.state_dict(), device=cuda, dtype=dtype, block_out_channels=G.INV_UNET["block_out_channels"])
    fp = "fp16" if dtype == torch.float16 else "bf16"
    I = Inverter(types.SimpleNamespace(unet=unet), DDIMSchedulerB200(), _cfg(fp=fp))
    got = I.ddim_inversion(x.to(cuda).to(dtype), conds.to(cuda).to(dtype))
    rel = ((got.float().cpu() - want).norm() / want.norm()).item()
    assert rel < tol, rel
    got0 = I.ddim_sample(want.to(cuda).to(dtype), conds.to(cuda).to(dtype))
    rel0 = ((got0.float().cpu() - want0).norm() / want0.norm()).item()
    assert rel0 < tol, rel0
