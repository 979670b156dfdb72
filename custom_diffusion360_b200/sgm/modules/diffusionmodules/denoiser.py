"""Denoiser preconditioning surface (reference: denoiser.py:7-79).

`denoiser(network, input, sigma, cond, **kwargs)` -> (D(x), fg_masks, alphas, rgbs), with
`.w(sigma)`, `.sigmas`, `sigma_to_idx`, `idx_to_sigma` as in the reference.  This generic entry
point evaluates the handful of scalar coefficient ops with torch; the sampling loop does not go
through it per step — it uses `FusedGuidedStep` (sampling.py), where c_in is folded into the input
convolution's im2col and c_out / c_skip / CFG / Euler are one kernel (cd360_cfg_euler_step).
"""
import torch
import torch.nn as nn

from ...util import append_dims, instantiate_from_config


class Denoiser(nn.Module):
    def __init__(self, weighting_config, scaling_config):
        super().__init__()
        self.weighting = instantiate_from_config(weighting_config)
        self.scaling = instantiate_from_config(scaling_config)

    def possibly_quantize_sigma(self, sigma):
        return sigma

    def possibly_quantize_c_noise(self, c_noise):
        return c_noise

    def w(self, sigma):
        return self.weighting(sigma)

    def __call__(self, network, input, sigma, cond, sigmas_ref=None, **kwargs):
        """Reference call (denoiser.py:22-44), including its training-time form
        `denoiser(network, x, σ, cond, input_ref=, sigmas_ref=, pose=, mask_ref=)` — what the reference's
        own loss passes (loss.py:171): second noising of the reference latents (:26-33), `c_in(σ_ref)`
        (:35-38), quantised `sigmas_ref`.  This entry point returns VALUES (there is no autograd in this
        build); the step that also produces gradients is `train_forward` + `UNetModel.backward`, which
        `StandardDiffusionLossImgRef` drives.  `noise_ref2` (optional kwarg) injects the second noise
        draw for parity tests."""
        sigma = self.possibly_quantize_sigma(sigma)
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        noise_ref2 = kwargs.pop("noise_ref2", None)
        if sigmas_ref is not None:
            kwargs["sigmas_ref"] = sigmas_ref
            if kwargs.get("input_ref") is not None:
                xr = kwargs["input_ref"]
                n2 = noise_ref2.to(xr) if noise_ref2 is not None else torch.randn_like(xr)
                kwargs["input_ref"] = xr + n2 * append_dims(sigmas_ref, xr.ndim)
        if kwargs.get("input_ref") is not None and "sigmas_ref" in kwargs:
            _, _, c_in_ref, _ = self.scaling(append_dims(kwargs["sigmas_ref"], kwargs["input_ref"].ndim))
            kwargs["input_ref"] = kwargs["input_ref"] * c_in_ref
            kwargs["sigmas_ref"] = self.possibly_quantize_c_noise(kwargs["sigmas_ref"])
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma_shape))
        predict, fg, alphas, rgbs = network(input * c_in, c_noise, cond, **kwargs)
        return predict * c_out + input * c_skip, fg, alphas, rgbs

    def train_forward(self, network, input, sigma, cond, sigmas_ref=None, input_ref=None, pose=None,
                      noise_ref2=None, jitter=None, mask_ref=None):
        """The training-time call (denoiser.py:22-44) up to the network output: quantise σ, noise the
        reference latents a second time (:26-33 — the loss already noised them once, loss.py:163-170;
        kept as in the reference), scale them by c_in(σ_ref) (:35-38), run the taped network with
        c_in folded into the input convolution's load.  Returns (eps fp32 tokens [b*hw, 4], aux,
        tape, quantised σ [b]); `D = eps c_out + input c_skip` is folded into the loss kernel."""
        sigma = self.possibly_quantize_sigma(sigma)
        kwargs = dict(pose=pose, jitter=jitter, mask_ref=mask_ref)
        if sigmas_ref is not None and input_ref is not None:
            n2 = noise_ref2.to(input_ref.device) if noise_ref2 is not None else torch.randn_like(input_ref)
            input_ref = input_ref + n2 * append_dims(sigmas_ref, input_ref.ndim)
            _, _, c_in_ref, _ = self.scaling(append_dims(sigmas_ref, input_ref.ndim))
            kwargs["input_ref"] = (input_ref * c_in_ref).float().contiguous()
            kwargs["sigmas_ref"] = self.possibly_quantize_c_noise(sigmas_ref)
        elif input_ref is not None:
            kwargs["input_ref"] = input_ref.float().contiguous()
        _, _, c_in, c_noise = self.scaling(sigma)
        c_noise = self.possibly_quantize_c_noise(c_noise)
        eps_tok, aux, tape = network.train_forward(input, c_noise, cond, in_scale=c_in.float().contiguous(),
                                                   **kwargs)
        return eps_tok, aux, tape, sigma


class DiscreteDenoiser(Denoiser):
    def __init__(self, weighting_config, scaling_config, num_idx, discretization_config,
                 do_append_zero=False, quantize_c_noise=True, flip=True):
        super().__init__(weighting_config, scaling_config)
        sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero,
                                                                flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise

    def sigma_to_idx(self, sigma):
        return (sigma - self.sigmas[:, None]).abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def possibly_quantize_sigma(self, sigma):
        return self.idx_to_sigma(self.sigma_to_idx(sigma))

    def possibly_quantize_c_noise(self, c_noise):
        return self.sigma_to_idx(c_noise) if self.quantize_c_noise else c_noise
