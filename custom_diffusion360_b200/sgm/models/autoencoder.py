"""First-stage decode on the sm_100a kernels.

Reference: sgm/models/autoencoder.py — `AutoencoderKL` (:282-316) / `AutoencoderKLInferenceWrapper`
(:319-321), the `first_stage_config.target` of configs/train_co3d_concept.yaml:98-117.  Same
constructor keys (`embed_dim`, `ddconfig`, `lossconfig`, `ckpt_path`, `monitor`, ...) and the same
state-dict keys for everything decode touches (`post_quant_conv.*`, `decoder.*`); `encoder.*` /
`quant_conv.*` entries of sdxl_vae.safetensors are accepted and ignored (`strict=False`).

Only `decode` is built: it runs once per image right after the sampling loop (sample.py:194).
`encode` belongs to the training data path (images -> latents, outside SURVEY §8) and raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ..modules.diffusionmodules.model import Decoder


class AutoencoderKL(nn.Module):
    def __init__(self, embed_dim: int, ddconfig=None, lossconfig=None, ckpt_path=None, ignore_keys=(),
                 monitor=None, input_key="jpg", **unused):
        super().__init__()
        assert ddconfig is not None and ddconfig["double_z"]
        self.embed_dim = embed_dim
        self.input_key = input_key
        self.monitor = monitor
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=()):
        """sgm/models/autoencoder.py:60-79: .safetensors or a torch checkpoint with `state_dict`."""
        if str(path).endswith("safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            sd = torch.load(path, map_location="cpu")["state_dict"]
        sd = {k: v for k, v in sd.items() if not any(k.startswith(ik) for ik in ignore_keys)}
        return self.load_decode_state_dict(sd)

    def load_decode_state_dict(self, sd: dict):
        """Load the decode-side entries of a full autoencoder state dict; returns (missing, ignored)."""
        own = self.state_dict()
        use = {k: v for k, v in sd.items() if k in own}
        ignored = [k for k in sd if k not in own]
        bad = [k for k in ignored if not k.startswith(("encoder.", "quant_conv.", "loss.", "regularization."))]
        if bad:
            raise KeyError(f"unexpected keys in the autoencoder checkpoint: {bad[:5]}")
        missing, _ = self.load_state_dict(use, strict=False)
        return missing, ignored

    def encode(self, x):
        raise NotImplementedError("VAE encode is on the training data path, outside SURVEY §8 / §8f")

    @torch.no_grad()
    def decode(self, z, scale: float = 1.0, **decoder_kwargs):
        """z fp32 [B, embed_dim, h, w] -> image fp32 [B, 3, 8h, 8w] (reference :313-316).  `scale`
        multiplies z on load (decode_first_stage's 1/scale_factor, folded into post_quant_conv)."""
        w = self.post_quant_conv.weight.detach().float().reshape(self.post_quant_conv.out_channels, -1).contiguous()
        zq = ops.pointwise_conv_nchw(z.float().contiguous(), w, self.post_quant_conv.bias.detach().float().contiguous(),
                                     scale=scale)
        return self.decoder(zq, **decoder_kwargs)


class AutoencoderKLInferenceWrapper(AutoencoderKL):
    pass
