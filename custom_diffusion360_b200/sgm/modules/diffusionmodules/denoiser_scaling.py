"""Preconditioning coefficients (reference: denoiser_scaling.py).  Only the ε-parameterisation the
shipped config uses (train_co3d_concept.yaml:23) is provided."""
import torch


class EpsScaling:
    def __call__(self, sigma):
        c_skip = torch.ones_like(sigma)
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = sigma.clone()
        return c_skip, c_out, c_in, c_noise
