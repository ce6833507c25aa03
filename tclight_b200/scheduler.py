"""DPM-Solver++ (SDE, Karras sigmas, 2nd-order midpoint) scheduler behind the interface TC-Light
uses (``set_timesteps``, ``timesteps``, ``step(eps, t, x, generator=..., return_dict=False)``,
``init_noise_sigma``) — the drop-in for diffusers' ``DPMSolverMultistepScheduler`` as constructed at
reference utils/model_utils.py:71-78.  The schedule and the scalar coefficients are computed on
the host with torch float32 0-dim arithmetic in the library's expression order (so they round
like the reference's); the tensor update runs in one CUDA kernel (tcl_dpm_step).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers randn_tensor semantics: a *list* of generators draws one (1, ...) sample per
    batch element, in order (reference generate.py:235 passes ``self.rng``, a list of N aliases of
    ONE generator, generate.py:568)."""
    if isinstance(generator, list):
        shape1 = (1,) + tuple(shape[1:])
        gdev = generator[0].device
        lat = [torch.randn(shape1, generator=generator[i], device=gdev, dtype=dtype) for i in range(shape[0])]
        return torch.cat(lat, dim=0).to(device)
    gdev = device if generator is None else generator.device
    return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)


class DPMSolverMultistepSchedulerB200:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, algorithm_type="sde-dpmsolver++",
                 use_karras_sigmas=True, steps_offset=1):
        if algorithm_type != "sde-dpmsolver++" or not use_karras_sigmas:
            raise NotImplementedError("only the configuration TC-Light constructs is implemented")
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)  # "linear" default
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.timesteps = None
        self.sigmas = None
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _sigma_to_t(sigma, log_sigmas):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        return ((1 - w) * low_idx + w * high_idx).reshape(sigma.shape)

    def set_timesteps(self, num_inference_steps, device=None):
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        log_sigmas = np.log(sigmas)
        sigmas = np.flip(sigmas).copy()
        smin, smax, rho = sigmas[-1].item(), sigmas[0].item(), 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        sigmas = (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho
        timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = [int(t) for t in timesteps]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _init_step_index(self, timestep):
        t = int(timestep.item()) if torch.is_tensor(timestep) else int(timestep)
        cand = [i for i, v in enumerate(self._timesteps_host) if v == t]
        if not cand:
            self._step_index = len(self._timesteps_host) - 1
        else:
            self._step_index = cand[1] if len(cand) > 1 else cand[0]

    def coefficients(self, i: int) -> dict:
        """fp32 scalars of step i, evaluated like the library does (0-dim float32 tensors)."""
        last = i == len(self._timesteps_host) - 1
        a_c, s_c = self._alpha_sigma(self.sigmas[i])
        a_n, s_n = self._alpha_sigma(self.sigmas[i + 1])
        lam_n = torch.log(a_n) - torch.log(s_n)
        lam_c = torch.log(a_c) - torch.log(s_c)
        h = lam_n - lam_c
        A = s_n / s_c * torch.exp(-h)
        B = a_n * (1 - torch.exp(-2.0 * h))
        Cn = s_n * torch.sqrt(1.0 - torch.exp(-2.0 * h))
        second = not (self.lower_order_nums < 1 or last)
        inv_r0 = 0.0
        if second:
            a_p, s_p = self._alpha_sigma(self.sigmas[i - 1])
            lam_p = torch.log(a_p) - torch.log(s_p)
            r0 = (lam_c - lam_p) / h
            inv_r0 = float(1.0 / r0)
        return dict(sigma_c_hat=float(s_c), alpha_c_hat=float(a_c), A=float(A), B=float(B), Cn=float(Cn),
                    inv_r0=inv_r0, second_order=second)

    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        if self._step_index is None:
            self._init_step_index(timestep)
        coef = self.coefficients(self._step_index)
        z = randn_tensor(model_output.shape, generator=generator, device=model_output.device, dtype=torch.float32)
        x0, x_next = ops.dpm_step(model_output.contiguous(), sample.contiguous(), self.model_outputs[1], z, coef)
        self.model_outputs[0] = self.model_outputs[1]
        self.model_outputs[1] = x0
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self._step_index += 1
        return (x_next,)

    def scale_model_input(self, sample, *a, **k):
        return sample


class DDIMSchedulerB200:
    """The part of diffusers' ``DDIMScheduler`` the reference's ``Inverter`` reads (invert.py:56-59, 151-244):
    ``set_timesteps`` / ``timesteps`` / ``alphas_cumprod`` / ``final_alpha_cumprod``, for the SD-1.5 scheduler
    config loaded at utils/VidToMe/utils.py:40-41 (scaled-linear betas 0.00085..0.012, 1000 train steps,
    ``timestep_spacing="leading"``, ``steps_offset=1``, ``set_alpha_to_one=False``).  The update itself is
    the Inverter's closed form (tcl_ddim_next), not ``scheduler.step``."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 set_alpha_to_one=False, steps_offset=1, timestep_spacing="leading"):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        if timestep_spacing != "leading":
            raise NotImplementedError("only the 'leading' spacing of the SD-1.5 scheduler config is implemented")
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, *a, **k):
        return sample
