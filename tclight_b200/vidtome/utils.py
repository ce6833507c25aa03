"""Mirror of the reference's utils/VidToMe/vidtome/utils.py:18-60 (RNG fork + frame reshapes)."""
import torch


def isinstance_str(x: object, cls_name: str) -> bool:
    """reference vidtome/utils.py:4-16."""
    return any(c.__name__ == cls_name for c in x.__class__.__mro__)


def init_generator(device: torch.device, fallback: torch.Generator = None) -> torch.Generator:
    """Forks the current default generator of `device` (reference vidtome/utils.py:18-30)."""
    device = torch.device(device)
    if device.type == "cpu":
        return torch.Generator(device="cpu").set_state(torch.get_rng_state())
    if device.type == "cuda":
        return torch.Generator(device=device).set_state(torch.cuda.get_rng_state())
    return init_generator(torch.device("cpu")) if fallback is None else fallback


def join_frame(x: torch.Tensor, fsize: int) -> torch.Tensor:
    """"(B F) N C -> B (F N) C" (vidtome/utils.py:32-35); a view for contiguous input."""
    BF, N, C = x.shape
    return x.reshape(BF // fsize, fsize * N, C)


def split_frame(x: torch.Tensor, fsize: int) -> torch.Tensor:
    """"B (F N) C -> (B F) N C" (vidtome/utils.py:37-40)."""
    B, FN, C = x.shape
    return x.reshape(B * fsize, FN // fsize, C)


def func_warper(funcs):
    def fn(x, **kw):
        for f in funcs:
            x = f(x, **kw)
        return x
    return fn


def join_warper(fsize):
    return lambda x, **kw: join_frame(x, fsize)


def split_warper(fsize):
    return lambda x, **kw: split_frame(x, fsize)
