"""σ schedules (reference: sgm/modules/diffusionmodules/discretizer.py:17-69, util.py:19-32)."""
from __future__ import annotations

import numpy as np
import torch


class Discretization:
    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sigmas = self.get_sigmas(n, device=device)
        if do_append_zero:
            sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
        return torch.flip(sigmas, (0,)) if flip else sigmas

    def get_sigmas(self, n, device):
        raise NotImplementedError


class LegacyDDPMDiscretization(Discretization):
    """Linear-in-sqrt(beta) DDPM schedule, float64 table; σ_t = sqrt((1 - ᾱ_t) / ᾱ_t).
    Sub-sampling picks `linspace(T-1, 0, n, endpoint=False).astype(int)[::-1]` like the reference.
    The table is always built on the host (float64 torch.linspace, as the reference does), whatever
    default device is active."""

    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps,
                                dtype=torch.float64, device="cpu") ** 2).numpy()
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)

    def get_sigmas(self, n, device="cpu"):
        if n < self.num_timesteps:
            steps = np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
            ac = self.alphas_cumprod[steps]
        elif n == self.num_timesteps:
            ac = self.alphas_cumprod
        else:
            raise ValueError(f"n={n} exceeds the {self.num_timesteps}-entry table")
        sigmas = torch.tensor((1 - ac) / ac, dtype=torch.float32, device="cpu") ** 0.5
        return torch.flip(sigmas, (0,)).to(device)
