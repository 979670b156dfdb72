"""Training loss of the pose-conditioned model (reference: StandardDiffusionLossImgRef,
sgm/modules/diffusionmodules/loss.py:95-216, 'l2' branch — the shipped config).

Same constructor keys and `__call__` signature as the reference; returns the reference's 4-tuple
`(loss_l2 [b], loss_fg [b,K], loss_bg [b,K], loss_rgb [b,K])` (K = pose blocks).  There is no
autograd: the call keeps the taped UNet forward, and `backward(...)` — called by
`DiffusionEngine.forward` once the weights of the terms in the total loss are known
(diffusion.py:221-236) — seeds the gradient kernels and runs the UNet backward.

Random draws (σ indices, the three noise tensors, stratified-sampling variates) are taken from
`batch["rand"]` when present (keys: sigma_idx | sigma, sigma_ref_idx | sigma_ref, noise, noise_ref,
noise_ref2, jitter),
so a test can replay exactly what the oracle / reference drew; otherwise torch's generator is used
like the reference does.
"""
from __future__ import annotations

import math
from types import SimpleNamespace as NS
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .... import ops
from ...util import append_dims, instantiate_from_config


class StandardDiffusionLossImgRef(nn.Module):
    def __init__(self, sigma_sampler_config, sigma_sampler_config_ref=None, type="l2",
                 offset_noise_level=0.0, batch2model_keys: Optional[Union[str, List[str]]] = None):
        super().__init__()
        if type != "l2":
            raise NotImplementedError("only the 'l2' loss of the shipped config is built")
        if offset_noise_level != 0.0:
            raise NotImplementedError("offset_noise_level is 0 in the shipped config")
        self.sigma_sampler = instantiate_from_config(sigma_sampler_config)
        self.sigma_sampler_ref = (instantiate_from_config(sigma_sampler_config_ref)
                                  if sigma_sampler_config_ref is not None else None)
        self.type = type
        self.offset_noise_level = offset_noise_level
        if not batch2model_keys:
            batch2model_keys = []
        if isinstance(batch2model_keys, str):
            batch2model_keys = [batch2model_keys]
        self.batch2model_keys = set(batch2model_keys)
        self.last: Optional[NS] = None

    # ---- forward: noising, denoiser / UNet (taped), loss values -------------------------------------
    def __call__(self, network, denoiser, conditioner, input, input_rgb, input_ref, pose, mask, mask_ref,
                 opacity, batch):
        cond = conditioner(batch) if conditioner is not None else batch["cond"]
        rnd = batch.get("rand", {}) if isinstance(batch, dict) else {}
        dev = input.device
        b = input.shape[0]
        sigmas = rnd["sigma"] if "sigma" in rnd else self.sigma_sampler(b, rand=rnd.get("sigma_idx")).to(dev)
        noise = rnd["noise"].to(dev) if "noise" in rnd else torch.randn_like(input)
        noised_input = (input + noise * append_dims(sigmas, input.ndim)).float().contiguous()
        extra = {}
        sigmas_ref = None
        if self.sigma_sampler_ref is not None:
            sigmas_ref = (rnd["sigma_ref"] if "sigma_ref" in rnd
                          else self.sigma_sampler_ref(b, rand=rnd.get("sigma_ref_idx")).to(dev))
            if input_ref is not None:
                nr = rnd["noise_ref"].to(dev) if "noise_ref" in rnd else torch.randn_like(input_ref)
                input_ref = input_ref + nr * append_dims(sigmas_ref, input_ref.ndim)   # loss.py:163-170
        self.last_cond = cond
        eps_tok, aux, tape, sigma_q = denoiser.train_forward(
            network, noised_input, sigmas, cond, sigmas_ref=sigmas_ref, input_ref=input_ref, pose=pose,
            noise_ref2=rnd.get("noise_ref2"), jitter=rnd.get("jitter"), mask_ref=mask_ref)
        self.last = NS(network=network, eps=eps_tok, aux=aux, tape=tape, sigma=sigma_q.float().contiguous(),
                       x_noisy=noised_input, target=input.float().contiguous(),
                       mask=None if mask is None else mask.float().contiguous())
        return self.get_loss(input_rgb, opacity)

    def _supervision_maps(self, input_rgb, opacity):
        """Per pose block: (opacity, mask, rgb target) resized like get_loss does (loss.py:183-206) —
        including the reference's quirk that `opacity` is re-assigned inside the loop, i.e. each
        block resizes the PREVIOUS block's map, not the original."""
        st = self.last
        maps = []
        op = None if opacity is None else opacity.float().contiguous()
        for _, (fg, _, _) in st.aux:
            size = int(math.sqrt(fg.shape[1]))
            if op is not None:
                op = ops.resize_bilinear_aa(op, size, size)
            mk = tg = None
            if st.mask is not None and input_rgb is not None:
                mk = ops.resize_bilinear_aa(st.mask, size, size)
                tg = ops.resize_bilinear_aa(input_rgb.float().contiguous(), size, size, scale=0.5, shift=0.5)
            maps.append((op, mk, tg))
        return maps

    def get_loss(self, input_rgb, opacity):
        st = self.last
        b = st.target.shape[0]
        dev = st.eps.device
        loss_l2, msum, _ = ops.diffusion_loss(st.eps, st.x_noisy, st.target, st.sigma, st.mask, 0.0)
        st.mask_sum = msum
        st.maps = self._supervision_maps(input_rgb, opacity) if opacity is not None else []
        zero = torch.zeros(b, device=dev)
        fg_l, bg_l, rgb_l = [], [], []
        for (blk, (fg, alphas, rgb)), (op, mk, tg) in zip(st.aux, st.maps):
            use_rgb = mk is not None
            loss3, _, _, _ = ops.nerf_aux_loss(fg, alphas, rgb if use_rgb else None, op.reshape(b, -1),
                                               mk.reshape(b, -1) if use_rgb else None,
                                               tg.reshape(b, 3, -1) if use_rgb else None, msum, zero, zero, zero)
            fg_l.append(loss3[:, 0])
            bg_l.append(loss3[:, 1])
            if use_rgb:
                rgb_l.append(loss3[:, 2])
        stack = lambda l: torch.stack(l, 1) if l else []
        return loss_l2, stack(fg_l), stack(bg_l), stack(rgb_l)

    # ---- backward: gradient seeds with the engine's weights, then the UNet backward ------------------
    def backward(self, coef_l2: float, w_fg=None, w_bg=None, w_rgb=None):
        """d(total)/d(pose weights) for total = coef_l2 * sum_b loss_l2[b] + sum_{b,k} (w_fg[b]
        loss_fg[b,k] + w_bg[b] loss_bg[b,k] + w_rgb[b] loss_rgb[b,k]); w_* fp32 [b] or None (term
        absent from the total)."""
        st = self.last
        b = st.target.shape[0]
        dev = st.eps.device
        _, _, deps = ops.diffusion_loss(st.eps, st.x_noisy, st.target, st.sigma, st.mask, coef_l2)
        daux_of = {}
        if st.maps and (w_fg is not None or w_bg is not None or w_rgb is not None):
            zero = torch.zeros(b, device=dev)
            wf = zero if w_fg is None else w_fg.float().contiguous()
            wb = zero if w_bg is None else w_bg.float().contiguous()
            wr = zero if w_rgb is None else w_rgb.float().contiguous()
            for (blk, (fg, alphas, rgb)), (op, mk, tg) in zip(st.aux, st.maps):
                use_rgb = mk is not None
                _, dfg, dal, drgb = ops.nerf_aux_loss(fg, alphas, rgb if use_rgb else None, op.reshape(b, -1),
                                                      mk.reshape(b, -1) if use_rgb else None,
                                                      tg.reshape(b, 3, -1) if use_rgb else None, st.mask_sum,
                                                      wf, wb, wr)
                daux_of[id(blk)] = (dfg, dal, drgb)
        cond_grads = st.network.diffusion_model.backward(st.tape, deps, daux_of)
        self.last = None  # free the tape
        return cond_grads
