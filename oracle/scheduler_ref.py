"""TEST INFRASTRUCTURE — restatement of diffusers==0.32.1 ``DPMSolverMultistepScheduler`` exactly as
TC-Light constructs it (reference utils/model_utils.py:71-78: 1000 train steps, linear betas
0.00085..0.012, ``sde-dpmsolver++``, Karras sigmas, order 2 midpoint, lower_order_final,
final sigma 0) and ``randn_tensor`` with a list of generators (reference generate.py:235, 568).

PARITY UNPINNED: diffusers is not vendored in /root/reference; this follows the published
algorithm.  The tensor-op order (and therefore the fp16 roundings of
``convert_model_output``) follows the library's expressions.
"""
from __future__ import annotations

import numpy as np
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: a generator *list* draws one sample at a time."""
    if isinstance(generator, list):
        shape1 = (1,) + tuple(shape[1:])
        gdev = generator[0].device
        lat = [torch.randn(shape1, generator=generator[i], device=gdev, dtype=dtype) for i in range(shape[0])]
        return torch.cat(lat, dim=0).to(device)
    gdev = device if generator is None else generator.device
    return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)


class DPMSolverSDEKarras:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.timesteps = None
        self.sigmas = None
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    # -- schedule -------------------------------------------------------------------------
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
        sigma_min, sigma_max = sigmas[-1].item(), sigmas[0].item()
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        sigmas = (sigma_max ** (1 / rho) + ramp * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
        timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas]).round()
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        cand = (self.timesteps == timestep).nonzero()
        if len(cand) == 0:
            self._step_index = len(self.timesteps) - 1
        elif len(cand) > 1:
            self._step_index = cand[1].item()
        else:
            self._step_index = cand[0].item()

    # -- one step -------------------------------------------------------------------------
    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        last = i == len(self.timesteps) - 1
        # epsilon -> x0 in the input dtype
        a_c, s_c = self._alpha_sigma(self.sigmas[i])
        x0 = (sample - s_c * model_output) / a_c
        self.model_outputs[0] = self.model_outputs[1]
        self.model_outputs[1] = x0
        sample = sample.to(torch.float32)
        noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device, dtype=torch.float32)
        a_n, s_n = self._alpha_sigma(self.sigmas[i + 1])
        lam_n = torch.log(a_n) - torch.log(s_n)
        lam_c = torch.log(a_c) - torch.log(s_c)
        h = lam_n - lam_c
        if self.lower_order_nums < 1 or last:
            x_t = ((s_n / s_c * torch.exp(-h)) * sample
                   + (a_n * (1 - torch.exp(-2.0 * h))) * x0
                   + s_n * torch.sqrt(1.0 - torch.exp(-2.0 * h)) * noise)
        else:
            a_p, s_p = self._alpha_sigma(self.sigmas[i - 1])
            lam_p = torch.log(a_p) - torch.log(s_p)
            h_0 = lam_c - lam_p
            r0 = h_0 / h
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            x_t = ((s_n / s_c * torch.exp(-h)) * sample
                   + (a_n * (1 - torch.exp(-2.0 * h))) * D0
                   + 0.5 * (a_n * (1 - torch.exp(-2.0 * h))) * D1
                   + s_n * torch.sqrt(1.0 - torch.exp(-2.0 * h)) * noise)
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self._step_index += 1
        return (x_t.to(model_output.dtype),)

    def scale_model_input(self, sample, *a, **k):
        return sample


class DDIMRef:
    """diffusers 0.32.1 DDIMScheduler, the attributes Inverter reads (reference invert.py:56-59, 219-233), for the
    SD-1.5 scheduler config (scaled_linear 0.00085..0.012, leading spacing, steps_offset 1, set_alpha_to_one
    False).  Restated from the published algorithm; parity unpinned (diffusers is not installable here)."""

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.timesteps = torch.arange(999, -1, -1)

    def set_timesteps(self, n, device=None):
        ratio = 1000 // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + 1)
