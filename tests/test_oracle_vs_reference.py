"""Pins the oracle restatements against the UNMODIFIED reference code (imported through
oracle/refshim.py).  Runs only where /root/reference exists (the build container)."""
import copy

import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return refshim.import_reference()


class _Mod:
    pass


@pytest.mark.parametrize("F,n,C,hw", [(4, 64, 32, (8, 8)), (3, 64, 32, (8, 8)), (2, 256, 16, (16, 16)), (1, 64, 32, (8, 8))])
def test_compute_merge_matches_reference(ref, F, n, C, hw):
    from oracle import vidtome_ref as V

    torch.manual_seed(0)
    args = dict(max_downsample=2, generator=None, seed=123, batch_size=2, align_batch=True, merge_global=True,
                global_merge_ratio=0.5, local_merge_ratio=0.6, global_rand=0.5, target_stride=4)
    mod = _Mod()
    mod.generator = torch.Generator().manual_seed(7)
    state = V.MergeState(torch.Generator().manual_seed(7))
    info = dict(size=hw, args=copy.deepcopy(args))
    for it in range(4):   # several chunks: first seeds the pool, later ones merge against it
        x = torch.randn(2 * F, n, C)
        m, u, merged_ref = ref.patch.compute_merge(mod, x, info)
        merged, unmerge, trace = V.compute_merge(state, x, hw, args)
        assert torch.equal(merged, merged_ref), f"merged tokens differ at chunk {it}"
        y = torch.randn_like(merged)
        assert torch.equal(unmerge(y), u(y)), f"unmerge differs at chunk {it}"
        assert torch.equal(state.global_tokens, mod.global_tokens)


def test_compute_merge_skips_low_res(ref):
    from oracle import vidtome_ref as V

    args = dict(max_downsample=2, generator=None, seed=123, batch_size=2, align_batch=True, merge_global=True,
                global_merge_ratio=0.5, local_merge_ratio=0.6, global_rand=0.5, target_stride=4)
    x = torch.randn(8, 4, 16)          # 16x16 image seen at downsample 8
    state = V.MergeState(torch.Generator().manual_seed(1))
    merged, unmerge, _ = V.compute_merge(state, x, (16, 16), args)
    assert merged is x and state.global_tokens is None


def test_ddim_sample_oracle_matches_reference_generator(ref):
    """The restated sampler loop + oracle ToMe patch reproduce the reference's own
    Generator.ddim_sample / temporal_denoise / pred_noise (run unmodified) bit for bit on CPU."""
    import numpy as np
    from oracle import harness, pipeline_ref as P
    from oracle.unet_ref import make_unet

    kw = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
    gen_cfg = dict(n_timesteps=3, alpha_t=0.01, win_size_t=6)
    N, h, w = 8, 16, 16
    torch.manual_seed(3)
    x = torch.randn(1, 4, h, w).repeat(N, 1, 1, 1)
    cc = torch.randn(N, 4, h, w) * 0.18215
    conds = torch.randn(2, 20, 64)
    conds_t = torch.randn(2, 10, 64)

    def seed_all():
        torch.manual_seed(12345)
        np.random.seed(12345)

    u1 = make_unet(seed=0, **kw)
    g, _ = harness.make_reference_generator(unet=u1, gen=gen_cfg)
    seed_all()
    g.rng = [torch.Generator().manual_seed(12345)] * N
    want = g.ddim_sample(x.clone(), conds, conds_t, cc)

    u2 = make_unet(seed=0, **kw)
    seed_all()
    got = P.ddim_sample_oracle(u2, x.clone(), conds, conds_t, cc, n_timesteps=3, alpha_t=0.01, win_size_t=6,
                               rng=[torch.Generator().manual_seed(12345)] * N)
    assert torch.equal(got, want)


def _ref_generator_with_dataset(ref, edited, flows, masks, unq_inv, opt):
    from oracle import harness
    from oracle.unet_ref import make_unet

    tiny = make_unet(seed=0, block_out_channels=(64, 64, 64, 64), cross_attention_dim=64)
    g, _ = harness.make_reference_generator(unet=tiny, opt=opt)
    g.dataset = ref.dataloader.OptDataset(edited.clone(), flows.clone(), masks.clone(), device="cpu")
    g.data_parser.unq_inv = unq_inv.clone()
    return g


def test_stage2_oracle_matches_reference(ref):
    """oracle/postopt_ref.stage2_uvt == the reference's own unique_tensor_optimization on CPU."""
    from oracle import postopt_ref as O

    edited, flows, masks, inv = O.synthetic_clip(n=6, h=176, w=192, seed=1)
    opt = dict(epochs=2, batch_size=4)
    g = _ref_generator_with_dataset(ref, edited, flows, masks, inv, opt)
    torch.manual_seed(5)
    want_img, want_loss = g.unique_tensor_optimization()
    torch.manual_seed(5)
    batches = O.draw_batches(6, 4, 2)
    got_img, _, got_loss = O.stage2_uvt(edited, flows, masks, inv, batches)
    assert len(got_loss) == len(want_loss)
    assert max(abs(a - b) for a, b in zip(got_loss, want_loss)) < 1e-6
    assert (got_img - want_img).abs().max().item() < 1e-5


def test_stage1_oracle_matches_reference(ref, monkeypatch):
    """oracle stage1_exposure == the reference's exposure_align (its hard-coded device="cuda" at
    generate.py:378 is redirected to the CPU for this check only)."""
    from oracle import postopt_ref as O

    real_eye = torch.eye
    monkeypatch.setattr(ref.generate.torch, "eye", lambda *a, device=None, **k: real_eye(*a, **k))
    edited, flows, masks, inv = O.synthetic_clip(n=6, h=176, w=192, seed=2)
    opt = dict(epochs_exposure=2, batch_size=4)
    g = _ref_generator_with_dataset(ref, edited, flows, masks, inv, opt)
    torch.manual_seed(6)
    want_img, want_loss = g.exposure_align()
    torch.manual_seed(6)
    batches = O.draw_batches(6, 4, 2)
    got_img, _, got_loss = O.stage1_exposure(edited, flows, masks, batches)
    assert len(got_loss) == len(want_loss)
    assert max(abs(a - b) for a, b in zip(got_loss, want_loss)) < 1e-6
    assert (got_img - want_img).abs().max().item() < 1e-5


@pytest.mark.parametrize("seed,alpha,thr", [(0, 0.5, 0.05), (1, 0.1, 0.01), (2, 0.5, 0.2)])
def test_flowid_producer_matches_reference(ref, seed, alpha, thr):
    """get_soft_mask_bwds / get_flowid / voxelization / warp_flow (utils/flow_utils.py, general_utils.py:223) vs
    oracle/flowid_ref.py, including the last-writer-wins collision rule of the CPU index assignment."""
    from oracle import flowid_ref as R

    frames, fwd, bwd = R.synthetic_scene(n=5, h=36, w=44, seed=seed)
    m_ref = ref.flow_utils.get_soft_mask_bwds(frames * 2 - 1, fwd, bwd, alpha=alpha)
    m = R.soft_mask_bwds(frames * 2 - 1, fwd, bwd, alpha=alpha)
    assert torch.allclose(m, m_ref, rtol=0, atol=1e-6)
    ids_ref = ref.flow_utils.get_flowid(frames, fwd, m_ref, rgb_threshold=thr)
    ids = R.flow_ids(frames, fwd, m_ref, rgb_threshold=thr)
    assert torch.equal(ids, ids_ref)
    inv_ref = ref.general_utils.voxelization(ids_ref.view(-1, 1), frames.permute(0, 2, 3, 1).reshape(-1, 3), None, None)
    assert torch.equal(R.unique_inverse(ids), inv_ref.reshape(-1))
    # non-dense ids: the inverse is a true rank map, not the identity
    sparse = ids_ref.reshape(-1) * 3 + 7
    inv2 = ref.general_utils.voxelization(sparse.view(-1, 1), torch.zeros(sparse.numel(), 3), None, None)
    assert torch.equal(R.unique_inverse(sparse), inv2.reshape(-1))
    assert torch.equal(R.warp(frames, bwd), ref.flow_utils.warp_flow(frames, bwd))


def test_ddim_inversion_matches_reference_inverter(ref):
    from oracle import make_goldens as G, pipeline_ref as P
    from oracle.scheduler_ref import DDIMRef
    from oracle.unet_ref import make_unet

    gold = G.golden_inversion()
    x, conds = G.inversion_inputs()
    sch = DDIMRef()
    sch.set_timesteps(5)
    unet = make_unet(seed=0, **G.INV_UNET)
    with torch.no_grad():
        xT = P.ddim_walk(unet, sch, x, conds, 4, True)
        assert torch.equal(xT, gold["x_T"])
        assert torch.equal(P.ddim_walk(unet, sch, xT, conds, 4, False), gold["x_recon"])


def test_encode_prompt_pair_matches_reference(ref):
    """Generator.encode_prompt_inner / encode_prompt_pair (generate.py:97-135: chunking of long prompts into 75-token
    pieces) on a stand-in tokenizer / text encoder: the B200 mirror returns the reference's tensors."""
    import types

    from tclight_b200.generate import Generator as Mine

    class Tok:
        model_max_length, bos_token_id, eos_token_id = 77, 1, 2

        def __call__(self, txt, truncation=False, add_special_tokens=False):
            return {"input_ids": [3 + (hash(w) % 50) for w in txt.split()]}

    emb = torch.nn.Embedding(64, 16)
    enc = lambda ids: types.SimpleNamespace(last_hidden_state=emb(ids) + torch.arange(ids.shape[1])[None, :, None] * 0.01)
    g_ref = object.__new__(ref.generate.Generator)
    g_mine = object.__new__(Mine)
    for g in (g_ref, g_mine):
        torch.nn.Module.__init__(g)
        g.tokenizer, g.text_encoder, g.device = Tok(), enc, "cpu"
    short = "a sunlit kitchen warm light"
    long = " ".join(f"w{i}" for i in range(170))           # 3 chunks
    for pos, neg in [(short, "bad"), (long, short), (short, long), (long, long)]:
        c_ref, uc_ref = g_ref.encode_prompt_pair(pos, neg)
        c, uc = g_mine.encode_prompt_pair(pos, neg)
        assert c.shape == c_ref.shape and torch.equal(c, c_ref) and torch.equal(uc, uc_ref)


def test_oracle_unet_matches_reference_pnp_restatements(ref):
    """Pins the restated UNet arithmetic (oracle/unet_ref.py — diffusers is not vendored) to the reference's OWN
    restatements of it: ``register_attention_control`` re-implements Attention.forward (to_q/k/v, head split,
    QK^T * scale, softmax, PV, head merge, to_out) on the decoder attn1 modules and ``register_conv_control``
    re-implements ResnetBlock2D.forward on up_blocks[1].resnets[1] (utils/VidToMe/pnp_utils.py:40-172).  With an
    empty injection schedule they are pure re-statements; swapping them into the oracle UNet must not change its
    output.  The diffusers-API attributes those closures read are attached as thin adapters."""
    import importlib
    import os
    import types

    import torch.nn.functional as F
    from oracle.unet_ref import Attention, ResnetBlock2D, make_unet

    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        pnp = importlib.import_module("utils.VidToMe.pnp_utils")
    finally:
        os.chdir(cwd)

    kw = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
    unet = make_unet(seed=0, **kw)
    torch.manual_seed(11)
    Fr, h, w = 2, 12, 20                       # odd pyramid: exercises upsample_size
    x = torch.randn(2 * Fr, 4, h, w)
    cc = torch.randn(Fr, 4, h, w) * 0.2
    ehs = torch.randn(2 * Fr, 20, 64)
    t = torch.tensor(801)
    with torch.no_grad():
        want = unet(x, t, encoder_hidden_states=ehs, cross_attention_kwargs={"concat_conds": cc}).sample

    def head_to_batch_dim(self, tensor):       # diffusers Attention.head_to_batch_dim: [B, N, C] -> [B*h, N, C/h]
        B, N, C = tensor.shape
        return tensor.reshape(B, N, self.heads, C // self.heads).permute(0, 2, 1, 3).reshape(B * self.heads, N, C // self.heads)

    def batch_to_head_dim(self, tensor):       # inverse: [B*h, N, d] -> [B, N, h*d]
        Bh, N, d = tensor.shape
        return tensor.reshape(Bh // self.heads, self.heads, N, d).permute(0, 2, 1, 3).reshape(Bh // self.heads, N, self.heads * d)

    n_attn = n_res = 0
    for m in unet.modules():
        if isinstance(m, Attention):
            m.head_to_batch_dim = types.MethodType(head_to_batch_dim, m)
            m.batch_to_head_dim = types.MethodType(batch_to_head_dim, m)
            m.scale = (m.to_q.out_features // m.heads) ** -0.5
            n_attn += 1
        if isinstance(m, ResnetBlock2D):
            m.nonlinearity = F.silu
            m.upsample = m.downsample = None
            m.time_embedding_norm = "default"
            m.dropout = torch.nn.Identity()
            m.output_scale_factor = 1.0
            n_res += 1
    assert n_attn == 32 and n_res == 22
    model = types.SimpleNamespace(unet=unet)
    pnp.register_attention_control(model, [], 2)      # replaces attn1.forward on 8 decoder blocks
    pnp.register_conv_control(model, [], 2)           # replaces up_blocks[1].resnets[1].forward
    pnp.register_time(model, 801)
    assert "forward" in unet.up_blocks[2].attentions[0].transformer_blocks[0].attn1.__dict__
    assert "forward" in unet.up_blocks[1].resnets[1].__dict__
    with torch.no_grad():
        got = unet(x, t, encoder_hidden_states=ehs, cross_attention_kwargs={"concat_conds": cc}).sample
    err = ((got - want).norm() / want.norm()).item()
    assert err < 2e-6, err                           # einsum+softmax vs SDPA: fp32 round-off only
