"""TEST INFRASTRUCTURE — not product code.

Stub-import harness that makes the *unmodified* reference (Linketic/TC-Light, mounted read-only
at /root/reference in the build container) importable without its absent third-party
dependencies, so that its own first-party functions can be executed on CPU to (a) validate the
restatements in ``oracle/`` and (b) generate the golden vectors committed under
``tests/golden/`` (see oracle/make_goldens.py).

Nothing here is used at run time on the GPU box (/root/reference does not exist there); only
oracle/make_goldens.py and the build-container-only tests import it.

What is stubbed: diffusers, omegaconf, controlnet_aux, clip, lpips, imageio,
skimage, av, cosmos1.*, matplotlib, xformers, peft; ``torchvision.io.read_video/write_video``;
functional stand-ins for ``torch_scatter.scatter`` (mean/sum) and
``pytorch_msssim.ssim.{gaussian_filter,_fspecial_gauss_1d}`` (restated from the published
pytorch-msssim algorithm: separable valid depthwise convolution, skipped when a dim < window).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("TCLIGHT_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "generate.py"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, name="stub"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{self._n}.{k}")

    def __call__(self, *a, **k):
        return _Anything(f"{self._n}()")

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-imports resolve

    def _getattr(k, _n=name):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{_n}.{k}")

    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    """Functional stand-in for torch_scatter.scatter along dim 0 (sum / mean)."""
    import torch

    assert dim == 0
    index = index.reshape(-1).long()
    n = int(index.max().item()) + 1 if dim_size is None else dim_size
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp_min(1)
        res = res / cnt.view(-1, *([1] * (src.dim() - 1)))
    elif reduce not in ("sum", "add"):
        raise NotImplementedError(reduce)
    return res


def _fspecial_gauss_1d(size, sigma):
    import torch

    coords = torch.arange(size, dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g.unsqueeze(0).unsqueeze(0)


def _gaussian_filter(inp, win):
    import torch.nn.functional as F

    assert all(ws == 1 for ws in win.shape[1:-1]), win.shape
    C = inp.shape[1]
    out = inp
    for i, s in enumerate(inp.shape[2:]):
        if s >= win.shape[-1]:
            out = F.conv2d(out, weight=win.transpose(2 + i, -1), stride=1, padding=0, groups=C)
    return out


_installed = False


def install() -> None:
    """Idempotently install the stubs and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT} (only exists in the build container)")
    for name in [
        "diffusers", "diffusers.utils", "diffusers.models", "diffusers.models.attention_processor",
        "controlnet_aux", "controlnet_aux.processor", "omegaconf", "clip", "lpips", "imageio",
        "skimage", "skimage.metrics", "av", "xformers", "peft", "matplotlib", "matplotlib.pyplot",
        "cosmos1", "cosmos1.models", "cosmos1.models.diffusion", "cosmos1.models.diffusion.prompt_upsampler",
        "cosmos1.models.diffusion.prompt_upsampler.video2world_prompt_upsampler_inference",
        "plyfile",
    ]:
        if name not in sys.modules:
            _stub_module(name)
    ts = _stub_module("torch_scatter")
    ts.scatter = _scatter
    pm = _stub_module("pytorch_msssim")
    pms = _stub_module("pytorch_msssim.ssim")
    pms.gaussian_filter = _gaussian_filter
    pms._fspecial_gauss_1d = _fspecial_gauss_1d
    pm.ssim = pms
    import torchvision.io as tio

    if not hasattr(tio, "read_video"):
        tio.read_video = None
    if not hasattr(tio, "write_video"):
        tio.write_video = None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def import_reference():
    """Returns a namespace with the reference's first-party modules on the two hot paths."""
    install()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        import importlib

        ns = types.SimpleNamespace()
        ns.merge = importlib.import_module("utils.VidToMe.vidtome.merge")
        ns.patch = importlib.import_module("utils.VidToMe.vidtome.patch")
        ns.vt_utils = importlib.import_module("utils.VidToMe.vidtome.utils")
        ns.loss_utils = importlib.import_module("utils.loss_utils")
        ns.flow_utils = importlib.import_module("utils.flow_utils")
        ns.sh_utils = importlib.import_module("utils.sh_utils")
        ns.general_utils = importlib.import_module("utils.general_utils")
        ns.dataloader = importlib.import_module("utils.dataloader")
        ns.generate_utils = importlib.import_module("utils.VidToMe.generate_utils")
        ns.generate = importlib.import_module("generate")
    finally:
        os.chdir(cwd)
    return ns
