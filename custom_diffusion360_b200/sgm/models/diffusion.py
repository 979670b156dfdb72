"""DiffusionEngine surface (reference: sgm/models/diffusion.py:43-557) around the B200 UNet.

Kept: constructor config keys, `.model` (OpenAIWrapper) / `.model.diffusion_model`, `.denoiser`,
`.sampler`, `sample(cond, uc, batch_size, num_steps, randn, shape, **kwargs)`,
`clear_rendered_feat()`, trainable-parameter selection by name (`trainkeys`).
Out of scope here: the text conditioner and the VAE first stage
are not instantiated — callers feed `cond` / `uc` embeddings and receive latents; Lightning is
not required (plain nn.Module).  Training (`training_step`) is a later row.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ..modules.diffusionmodules.sampling import FusedGuidedStep
from ..modules.diffusionmodules.wrappers import OPENAIUNETWRAPPER
from ..util import default, get_obj_from_str, instantiate_from_config


def _cfg_get(cfg, path, dflt=None):
    cur = cfg
    for key in path:
        try:
            cur = cur[key]
        except (KeyError, TypeError, IndexError):
            return dflt
    return cur


class DiffusionEngine(nn.Module):
    def __init__(self, network_config, denoiser_config, first_stage_config=None,
                 conditioner_config=None, sampler_config=None, optimizer_config=None,
                 scheduler_config=None, loss_fn_config=None, network_wrapper=None, ckpt_path=None,
                 use_ema=False, ema_decay_rate=0.9999, scale_factor=1.0,
                 disable_first_stage_autocast=False, input_key="jpg", log_keys=None,
                 no_cond_log=False, compile_model=False, trainkeys="pose", multiplier=0.05,
                 loss_rgb_lambda=20.0, loss_fg_lambda=10.0, loss_bg_lambda=20.0):
        super().__init__()
        if use_ema:
            raise NotImplementedError("use_ema=True is unused by the shipped config")
        self.log_keys = log_keys
        self.input_key = input_key
        self.trainkeys = trainkeys
        self.multiplier = multiplier
        self.loss_rgb_lambda, self.loss_fg_lambda, self.loss_bg_lambda = loss_rgb_lambda, loss_fg_lambda, loss_bg_lambda
        self.rgb = _cfg_get(network_config, ("params", "rgb"), False)
        self.rgb_predict = _cfg_get(network_config, ("params", "rgb_predict"), False)
        self.scale_factor = scale_factor
        self.use_ema = False
        model = instantiate_from_config(network_config)
        self.model = get_obj_from_str(default(network_wrapper, OPENAIUNETWRAPPER))(model, compile_model=compile_model)
        self.denoiser = instantiate_from_config(denoiser_config)
        self.sampler = instantiate_from_config(sampler_config) if sampler_config is not None else None
        # deliberately not built (out of the hot path): kept as configs for a caller that wants them
        self.conditioner_config = conditioner_config
        self.first_stage_config = first_stage_config
        self.loss_fn_config = loss_fn_config
        self.conditioner = None
        self.first_stage_model = None
        # trainable set by parameter name (reference :119-147)
        for name, p in self.model.diffusion_model.named_parameters():
            if trainkeys == "pose":
                p.requires_grad = "pose" in name
            elif trainkeys == "all":
                p.requires_grad = True
        self._fused: Optional[FusedGuidedStep] = None

    @property
    def device(self):
        return next(self.model.parameters()).device

    def clear_rendered_feat(self):
        self.model.diffusion_model.clear_rendered_feat()

    def set_reference_choices(self, choices):
        self.model.diffusion_model.set_reference_choices(choices)

    def decode_first_stage(self, z):
        raise NotImplementedError("VAE decode is the first 'next' row of SURVEY §8f; this engine returns latents")

    @torch.no_grad()
    def sample(self, cond: Dict, uc: Union[Dict, None] = None, batch_size: int = 16, num_steps=None,
               randn=None, shape: Union[None, Tuple, List] = None, return_rgb=False, mask=None,
               init_im=None, noise=None, fused: bool = True, **kwargs):
        """Reference contract (diffusion.py:375-401).  `noise` is accepted as an alias of `randn`
        (sample.py:190-192 passes `noise=`).  With `fused=True` (default) the loop runs the
        graph-replayed fused step; `fused=False` goes through the generic denoiser/sampler objects."""
        if mask is not None or init_im is not None:
            raise NotImplementedError("masked / img2img sampling is unused by sample.py")
        randn = noise if randn is None else randn
        if randn is None:
            randn = torch.randn(batch_size, *shape)
        x = randn.to(self.device).float().contiguous().clone()
        uc = default(uc, cond)
        if not fused:
            denoiser = lambda inp, sigma, c: self.denoiser(self.model, inp, sigma, c, **kwargs)
            samples, rgb_list = self.sampler(denoiser, x, cond, uc=uc, num_steps=num_steps)
            return (samples, rgb_list) if return_rgb else samples
        n_img = x.shape[0]
        pose = kwargs.get("pose")
        if isinstance(pose, (list, tuple)):
            pose = list(pose)[:n_img]  # sample.py passes `pose * rows`
        step = FusedGuidedStep(self.model.diffusion_model, self.denoiser, self.sampler.guider, cond, uc,
                               pose=pose, n_img=n_img, latent_shape=tuple(x.shape[1:]))
        self._fused = step
        samples = self.sampler.sample_fused(step, x, num_steps=num_steps)
        return (samples, None) if return_rgb else samples
