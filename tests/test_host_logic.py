"""Host-side mirrors (no GPU): chunk planning, window plan, LR schedule, config loader, scheduler
coefficients, optimiser batch order — against the oracle restatements (pinned to the reference)."""
import math
import os
import types

import numpy as np
import pytest
import torch


def _gen_stub(**kw):
    from tclight_b200.generate import VidToMeGenerator

    g = types.SimpleNamespace(chunk_size=4, merge_global=True, chunk_ord="mix", perm_div=4.0)
    g.__dict__.update(kw)
    g.get_chunks = types.MethodType(VidToMeGenerator.get_chunks, g)
    return g


@pytest.mark.parametrize("n", [1, 3, 8, 30, 160, 300])
@pytest.mark.parametrize("ord_", ["mix", "rand", "seq"])
def test_get_chunks_matches_oracle_and_partitions(n, ord_):
    from oracle.pipeline_ref import plan_chunks

    g = _gen_stub(chunk_ord=ord_)
    for seed in range(5):
        np.random.seed(seed); torch.manual_seed(seed)
        a = g.get_chunks(n)
        np.random.seed(seed); torch.manual_seed(seed)
        b = plan_chunks(n, 4, True, ord_, 4.0)
        assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
        allidx = torch.cat(a).sort().values
        assert torch.equal(allidx, torch.arange(n))                      # a partition of the frames
        assert all(1 <= len(c) <= 4 for c in a)
        assert all(int(c[-1]) - int(c[0]) + 1 == len(c) for c in a)      # consecutive runs


def test_temporal_window_plan():
    from oracle.pipeline_ref import window_plan
    from tclight_b200.generate import Generator

    assert Generator.temporal_windows(300, 64) == ([0, 59, 118, 177, 236], [5, 5, 5, 5])
    assert Generator.temporal_windows(30, 64) == ([0], [0])
    for n in (2, 63, 64, 65, 127, 128, 171, 250, 600):
        assert Generator.temporal_windows(n, 64) == window_plan(n, 64)
        starts, _ = Generator.temporal_windows(n, 64)
        assert starts[-1] + 64 >= n and all(b - a <= 64 for a, b in zip(starts, starts[1:]))


def test_expon_lr_indexing():
    from oracle.postopt_ref import expon_lr
    from tclight_b200.postopt import get_expon_lr_func

    N, Bo, epochs = 300, 16, 35
    total = epochs * N // Bo
    f = get_expon_lr_func(0.01, 0.001, lr_delay_steps=0, lr_delay_mult=0.0, max_steps=total)
    assert f(0) == pytest.approx(0.01) and f(total) == pytest.approx(0.001) and f(total + 5) == pytest.approx(0.001)
    for step in (1, 19, 20, 333, total):
        assert f(step) == pytest.approx(expon_lr(step, 0.01, 0.001, total), rel=1e-12)
    # reference quirk (generate.py:394): indices are not strictly monotonic across epochs
    idx = [ep * N // Bo + i + 1 for ep in range(2) for i in range(math.ceil(N / Bo))]
    assert idx[18] == 19 and idx[19] == 19


def test_config_loader_chain_and_interpolation(tmp_path):
    from tclight_b200 import config_utils as CU

    base = CU.DEFAULT_YAML
    child = tmp_path / "droid.yaml"
    child.write_text(f"work_dir: 'wd/x'\ndata:\n  rgb_path: 'examples/droid.mp4'\n  height: 536\ngeneration:\n  alpha_t: 0.01\n"
                     f"  frame_range: [0, -1, 1]\n  prompt:\n    droid: an office\nbase_config: {base}\n")
    cfg = CU.resolve(CU.load_config_file(str(child)))
    assert cfg.data.height == 536 and cfg.data.width == 960                 # child wins, base fills
    assert cfg.generation.alpha_t == 0.01 and cfg.generation.n_timesteps == 25
    assert cfg.inversion.save_path == "wd/x/latents" and cfg.generation.latents_path == "wd/x/latents"
    assert cfg.generation.output_path == "wd/x"
    assert cfg.post_opt.epochs == 70 and cfg.post_opt.batch_size == 16
    assert "float_precision" not in cfg.generation and cfg.float_precision == "fp16"
    cfg2 = CU.load_config(print_config=False, argv=["--config", str(child), "--multi_axis", "-n", "bad"])
    assert cfg2.generation.alpha_t == 0.01 and cfg2.generation.negative_prompt == "bad"
    assert dict(cfg2.generation.prompt) == {"droid": "an office"}
    CU.save_config(cfg2, str(tmp_path / "out"), gene=True)
    assert os.path.exists(tmp_path / "out" / "config.yaml")


def test_scheduler_coefficients_match_oracle_expressions():
    from oracle.scheduler_ref import DPMSolverSDEKarras
    from tclight_b200.scheduler import DPMSolverMultistepSchedulerB200

    ref, mine = DPMSolverSDEKarras(), DPMSolverMultistepSchedulerB200()
    ref.set_timesteps(25)
    mine.set_timesteps(25)
    assert torch.equal(ref.timesteps, mine.timesteps) and torch.equal(ref.sigmas, mine.sigmas)
    assert len(mine.timesteps) == 25 and float(mine.sigmas[-1]) == 0.0
    x = torch.randn(2, 4, 4, 4)
    for i in range(25):
        c = mine.coefficients(i)
        assert c["second_order"] == (0 < i < 24)
        eps = torch.randn(2, 4, 4, 4)
        x_ref = ref.step(eps, ref.timesteps[i], x, generator=torch.Generator().manual_seed(i))[0]
        # replay with the scalar coefficients (what the CUDA kernel evaluates)
        z = torch.randn(2, 4, 4, 4, generator=torch.Generator().manual_seed(i))
        x0 = (x - c["sigma_c_hat"] * eps) / c["alpha_c_hat"]
        acc = c["A"] * x + c["B"] * x0 + c["Cn"] * z
        if c["second_order"]:
            acc = acc + 0.5 * c["B"] * c["inv_r0"] * (x0 - prev_x0)
        assert torch.allclose(acc, x_ref, rtol=1e-5, atol=1e-5), i
        prev_x0 = x0
        mine.lower_order_nums = min(mine.lower_order_nums + 1, 2)
        x = x_ref
    assert torch.allclose(x, prev_x0, atol=1e-6)      # final sigma 0 => x' = x0


def test_batch_iterator_matches_dataloader_order():
    from oracle.postopt_ref import draw_batches
    from tclight_b200.postopt import batch_iterator

    torch.manual_seed(3)
    want = draw_batches(37, 16, 3)
    torch.manual_seed(3)
    loader = batch_iterator(37, 16)
    got = [[[int(i) for i in b] for b in loader] for _ in range(3)]
    assert got == want and [len(b) for b in got[0]] == [16, 16, 5]


def test_random_state_dict_has_diffusers_sd15_layout():
    from tclight_b200.weights import interleave_geglu, pack_conv3x3, random_state_dict

    sd = random_state_dict(block_out_channels=(64, 128, 256, 256))
    assert sd["conv_in.weight"].shape == (64, 8, 3, 3)
    assert sd["up_blocks.3.resnets.0.conv1.weight"].shape[1] == 128 + 64
    assert "down_blocks.3.attentions.0.norm.weight" not in sd and "mid_block.attentions.0.proj_in.weight" in sd
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv3x3(w)
    assert p.shape == (2, 27) and p[1, (1 * 3 + 2) * 3 + 0] == w[1, 0, 1, 2]
    wi, bi = interleave_geglu(torch.arange(512.)[:, None].repeat(1, 2), torch.arange(512.))
    assert wi[0, 0] == 0 and wi[128, 0] == 256 and wi[256, 0] == 128 and bi[384] == 384


def test_new_mirrors_refuse_cpu_tensors():
    """flow_utils / invert / vae are CUDA-only product paths (no CPU fallback)."""
    import types

    import pytest
    import torch
    from tclight_b200 import flow_utils
    from tclight_b200._lib import TclError
    from tclight_b200.invert import Inverter
    from tclight_b200.scheduler import DDIMSchedulerB200
    from tclight_b200.vae import AutoencoderKLB200

    x = torch.zeros(2, 3, 16, 16)
    f = torch.zeros(2, 2, 16, 16)
    with pytest.raises(TclError):
        flow_utils.warp_flow(x, f)
    with pytest.raises(TclError):
        flow_utils.get_soft_mask_bwds(x, f, f)
    with pytest.raises(TclError):
        flow_utils.get_flowid(x, f, torch.ones(2, 1, 16, 16))
    with pytest.raises(TclError):
        flow_utils.voxelization(torch.zeros(8, 1, dtype=torch.int32))
    with pytest.raises(TclError):
        AutoencoderKLB200({}, device="cpu")

    class D(dict):
        __getattr__ = dict.__getitem__

    cfg = types.SimpleNamespace(device="cuda", sd_version="1.5", model_key=None, float_precision="fp16", height=64, width=64,
                                inversion=D(control="none", control_scale=1.0, save_steps=5, steps=5, prompt="", recon=False,
                                            save_intermediate=False, use_blip=False, batch_size=8, force=True, n_frames=None))
    inv = Inverter(types.SimpleNamespace(unet=None), DDIMSchedulerB200(), cfg)
    assert list(inv.scheduler.timesteps) == [801, 601, 401, 201, 1]
    mu, sg, mu_p, sg_p = inv._coefs(801, 0, False)
    assert abs(mu * mu + sg * sg - 1) < 1e-6 and abs(mu_p * mu_p + sg_p * sg_p - 1) < 1e-6 and mu_p > mu
    with pytest.raises(TclError):
        inv.pred_next_x(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 801, 0)
    cfg.inversion["control"] = "depth"
    with pytest.raises(TclError):
        Inverter(types.SimpleNamespace(unet=None), DDIMSchedulerB200(), cfg)


def test_vae_oracle_shapes():
    import torch
    from oracle import vae_ref as V

    m = V.make_vae(block_out_channels=(64, 64, 128, 128))
    keys = set(m.state_dict().keys())
    for k in ("encoder.down_blocks.0.downsamplers.0.conv.weight", "decoder.up_blocks.2.upsamplers.0.conv.weight",
              "encoder.mid_block.attentions.0.to_q.weight", "decoder.mid_block.attentions.0.to_out.0.bias",
              "quant_conv.weight", "post_quant_conv.bias", "decoder.up_blocks.3.resnets.0.conv_shortcut.weight" if False else "decoder.conv_norm_out.weight"):
        assert k in keys, k
    x = torch.rand(1, 3, 40, 56)
    z = V.encode_imgs(m, x)
    assert z.shape == (1, 4, 5, 7)
    y = V.decode_latents(m, z)
    assert y.shape == x.shape and 0 <= float(y.min()) and float(y.max()) <= 1


def test_prefetch_draws_equals_sequential_draws():
    """vidtome.prefetch_draws consumes each block generator with the reference's calls in the reference's order
    (randint for the target frame when a chunk has > 1 frame, then rand for the global role once a pool exists), so the
    prefetched values are the ones the per-call draws would have produced."""
    import pytest
    import torch
    from tclight_b200._lib import TclError
    from tclight_b200.vidtome import patch

    args = dict(merge_global=True, local_merge_ratio=0.6, target_stride=4, global_rand=0.5, max_downsample=2)
    unet = torch.nn.Module()
    unet._tome_info = {"size": None, "args": args}
    blocks = []
    for k in range(3):
        b = torch.nn.Module()
        b._tome_info, b._tome_patched, b._tome_active = unet._tome_info, True, k != 2   # block 2: level skipped (ds > 2)
        b.generator = torch.Generator().manual_seed(100 + k)
        b.global_tokens = None if k == 0 else torch.zeros(1)                             # block 1 already has a pool
        setattr(unet, f"b{k}", b)
        blocks.append(b)
    fsizes = [3, 4, 4, 1, 4, 2]
    n = patch.prefetch_draws(unet, fsizes)
    assert not getattr(blocks[2], "_draw_queue", None)
    total = 0
    for k in (0, 1):
        ref = torch.Generator().manual_seed(100 + k)
        has_pool = k == 1
        want = []
        for fs in fsizes:
            if fs > 1:
                want.append(["i", min(4, fs), float(torch.randint(0, min(4, fs), (1,), generator=ref))])
            if has_pool:
                want.append(["r", 0, float(torch.rand(1, generator=ref))])
            has_pool = True
        assert [list(e) for e in blocks[k]._draw_queue] == want
        total += len(want)
    assert n == total
    with pytest.raises(TclError):
        patch.assert_draws_consumed(unet)
    with pytest.raises(TclError):
        patch.prefetch_draws(unet, [4])          # a pass must start with empty queues


def test_random_vae_state_dict_has_diffusers_keys():
    from oracle import vae_ref as V
    from tclight_b200.weights import random_vae_state_dict

    for boc in [(128, 256, 512, 512), (64, 64, 128, 128)]:
        sd = random_vae_state_dict(seed=1, block_out_channels=boc)
        ref = V.make_vae(block_out_channels=boc)
        assert set(sd) == set(ref.state_dict()) and all(sd[k].shape == v.shape for k, v in ref.state_dict().items())
        ref.load_state_dict(sd)


def test_video_dataparser_host_side(tmp_path):
    """Frame decoding / resizing / flow-cache layout of tclight_b200.dataparser (CPU-only parts)."""
    import cv2
    import numpy as np
    import pytest
    import torch
    import types
    from tclight_b200 import dataparser as D
    from tclight_b200._lib import TclError

    rng = np.random.default_rng(0)
    vdir = tmp_path / "clip"
    vdir.mkdir()
    raw = rng.integers(0, 255, size=(5, 40, 72, 3), dtype=np.uint8)
    for i, im in enumerate(raw):
        cv2.imwrite(str(vdir / f"{i:04d}.png"), cv2.cvtColor(im, cv2.COLOR_RGB2BGR))
    cfg = types.SimpleNamespace(rgb_path=str(vdir), height=32, width=48, fps=25)
    dp = D.VideoDataParser(cfg, device="cpu")
    assert dp.n_frames == 5 and dp.alpha == 0.5 and dp.flow_model == "memflow"
    rgbs = dp.load_video(frame_ids=[0, 2, 4])
    assert rgbs.shape == (3, 3, 32, 48) and 0 <= float(rgbs.min()) and float(rgbs.max()) <= 1
    # resize-to-cover + centre crop: 40x72 -> scale max(48/72, 32/40) = 0.8 -> 32x58 -> crop 32x48
    full = torch.from_numpy(raw[[0, 2, 4]]).permute(0, 3, 1, 2).float() / 255
    assert torch.equal(rgbs, D.process_frames(full, 32, 48))
    assert D.get_frame_ids([0, -1], 5) == [0, 1, 2, 3, 4] and D.get_frame_ids([1, 9], 5) == [1, 2, 3, 4]
    assert D.get_frame_ids([0, 3], 5, frame_ids=[4, 1]) == [1, 4]
    # flow cache layout of the reference: <frame dir>/future_flow_memflow/<id:04d>.pt, tensors [1,2,h,w] (the reference
    # estimates flows on the processed frames)
    fdir = vdir / "future_flow_memflow"
    fdir.mkdir()
    for i in (0, 2, 4):
        torch.save(torch.full((1, 2, 32, 48), float(i)), str(fdir / f"{i:04d}.pt"))
    flows, pasts, masks = dp.load_flow([0, 2, 4], future_flow=True, past_flow=False, gts=rgbs, save_flow=False)
    assert pasts is None and masks is None and flows.shape == (3, 2, 32, 48)
    assert torch.allclose(flows[1], torch.full((2, 32, 48), 2.0))
    # flows stored at another resolution are resized like the frames and their vectors scaled by the resize factor
    assert torch.allclose(dp.process_flow([torch.full((2, 40, 72), 5.0)]), torch.full((1, 2, 32, 48), 5.0 * 0.8))
    # a missing flow without a flow_fn is an error, not a silent zero
    (fdir / "0002.pt").unlink()
    with pytest.raises(TclError):
        dp.load_flow([0, 2, 4], future_flow=True, past_flow=False, gts=rgbs, save_flow=False)
    # ... and is produced (and cached) by a caller-supplied estimator
    dp.flow_fn = lambda s, t: torch.ones(1, 2, s.shape[-2], s.shape[-1])
    flows, _, _ = dp.load_flow([0, 2, 4], future_flow=True, past_flow=False, gts=rgbs, save_flow=True)
    assert (fdir / "0002.pt").exists() and torch.allclose(flows[1], torch.ones(2, 32, 48))


def test_process_frames_matches_reference():
    import pytest
    import torch
    from oracle import refshim

    if not refshim.reference_available():
        pytest.skip("reference tree not mounted")
    import importlib
    import os

    from tclight_b200 import dataparser as D

    refshim.install()
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        ru = importlib.import_module("utils.VidToMe.utils")
        rg = importlib.import_module("utils.general_utils")
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(0)
    fr = torch.rand(3, 3, 50, 90, generator=g)
    assert torch.equal(D.process_frames(fr, 32, 48), ru.process_frames(fr, 32, 48, 8))
    fl = torch.randn(3, 2, 50, 90, generator=g)
    assert torch.equal(D.process_frames(fl, 40, 64), rg.process_frames(fl, 40, 64))
    assert D.get_frame_ids([0, -1], 7) == ru.get_frame_ids([0, -1], 7)


def test_iclight_weight_surgery_matches_reference_recipe():
    """iclight_merge_state_dict == the reference's conv_in widening + offset addition (utils/model_utils.py:22-26, 50-54)."""
    import pytest
    import torch
    from tclight_b200._lib import TclError
    from tclight_b200.model_utils import iclight_merge_state_dict

    g = torch.Generator().manual_seed(0)
    origin = {"conv_in.weight": torch.randn(6, 4, 3, 3, generator=g), "conv_in.bias": torch.randn(6, generator=g),
              "mid.weight": torch.randn(5, 5, generator=g)}
    # the reference: new Conv2d(8, ...) with zero weight, copy the first 4 input channels, keep the bias; then add offsets
    conv = torch.nn.Conv2d(8, 6, 3, 1, 1)
    with torch.no_grad():
        conv.weight.zero_()
        conv.weight[:, :4].copy_(origin["conv_in.weight"])
    ref_origin = {"conv_in.weight": conv.weight.detach(), "conv_in.bias": origin["conv_in.bias"], "mid.weight": origin["mid.weight"]}
    offset = {k: torch.randn(v.shape, generator=g) for k, v in ref_origin.items()}
    want = {k: ref_origin[k] + offset[k] for k in ref_origin}
    got = iclight_merge_state_dict(origin, offset)
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    assert got["conv_in.weight"].shape == (6, 8, 3, 3) and torch.equal(got["conv_in.weight"][:, 4:], offset["conv_in.weight"][:, 4:])
    with pytest.raises(TclError):
        iclight_merge_state_dict(origin, {k: v for k, v in offset.items() if k != "mid.weight"})


def test_run_cli_requires_cuda_for_synthetic(capsys):
    import torch
    from tclight_b200 import run

    if torch.cuda.is_available():
        return
    assert run.main(["--synthetic", "--small"]) == 2
    assert "CUDA device is required" in capsys.readouterr().err


def test_save_video_and_frames_roundtrip(tmp_path):
    import cv2
    import torch
    from tclight_b200 import dataparser as D

    g = torch.Generator().manual_seed(0)
    frames = torch.rand(4, 3, 32, 48, generator=g)
    out = D.save_video(frames, str(tmp_path / "o"), save_frame=True, fps=10, post_fix="_x")
    assert out.endswith("output_x.mp4")
    cap = cv2.VideoCapture(out)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 4 and int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 48
    cap.release()
    back = D.read_frames(str(tmp_path / "o" / "frames_x"))            # PNGs are lossless: 8-bit truncation only
    assert back.shape == frames.shape and (back - frames).abs().max() <= 1 / 255 + 1e-6
    p = D.save_loss_curve([0.5, 0.25], str(tmp_path / "o"), "loss_exposure")
    assert open(p).read().split() == ["0.5", "0.25"]
