"""CPU oracle of the VAE decode that follows the sampling loop — TEST INFRASTRUCTURE ONLY.

SURVEY.md §8f row 1: `DiffusionEngine.decode_first_stage` (sgm/models/diffusion.py:207-212) ->
`AutoencoderKL.decode` (sgm/models/autoencoder.py:313-316) -> `Decoder.forward`
(sgm/modules/diffusionmodules/model.py:715-757) with `ResnetBlock` (:94-151), `Upsample` (:58-71),
`MemoryEfficientAttnBlock` (:204-266, single-head attention over all pixels) and `Normalize`
(:52-55, GroupNorm(32, eps 1e-6)).  Plain-torch fp32, functional (state dict in, tensors out), one
function per reference symbol.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.

Parity pinning: checked live against the reference's own `Decoder` imported in place
(tests/test_oracle_vs_reference.py::test_vae_decoder_oracle_vs_reference) and against golden vectors
generated from it (tests/golden/make_vae_golden.py -> vae_decoder_golden.pt).  xformers is not
installed: its `memory_efficient_attention` is stood in for by exact softmax attention
(oracle/ref_harness.py), as for the UNet.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# half-width copy of the shipped ddconfig (configs/train_co3d_concept.yaml:104-115: ch 128,
# ch_mult [1,2,4,4], 2 res blocks, no per-level attention); same topology, CPU-sized
TINY_VAE_CFG = dict(ch=64, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=(),
                    dropout=0.0, in_channels=3, resolution=256, z_channels=4, double_z=True,
                    attn_type="vanilla-xformers")
SDXL_VAE_CFG = dict(TINY_VAE_CFG, ch=128)
SDXL_SCALE_FACTOR = 0.13025   # configs/train_co3d_concept.yaml:7


def normalize(sd, p: str, x: Tensor) -> Tensor:
    """Normalize (model.py:52-55): GroupNorm(32, eps=1e-6, affine)."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def nonlinearity(x: Tensor) -> Tensor:
    """model.py:47-49 (swish)."""
    return x * torch.sigmoid(x)


def resnet_block(sd, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with temb=None (model.py:131-151); 1x1 `nin_shortcut` when the width changes."""
    h = F.conv2d(nonlinearity(normalize(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(nonlinearity(normalize(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def attn_block(sd, p: str, x: Tensor) -> Tensor:
    """MemoryEfficientAttnBlock.forward (model.py:231-266): 1x1 q/k/v, ONE head of width C over all
    H*W pixels, softmax(q k^T / sqrt(C)) v, 1x1 proj_out, residual."""
    h = normalize(sd, p + ".norm", x)
    q, k, v = (F.conv2d(h, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]) for n in "qkv")
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))      # b (h w) c
    w = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
    out = (w @ v).transpose(1, 2).reshape(b, c, hh, ww)
    return x + F.conv2d(out, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def upsample(sd, p: str, x: Tensor) -> Tensor:
    """Upsample.forward (model.py:67-71): nearest x2 then 3x3 conv."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)


def decoder(sd, cfg: dict, z: Tensor, prefix: str = "decoder") -> Tensor:
    """Decoder.forward (model.py:715-757), give_pre_end / tanh_out off."""
    p = prefix
    n_levels = len(cfg["ch_mult"])
    h = F.conv2d(z, sd[p + ".conv_in.weight"], sd[p + ".conv_in.bias"], padding=1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for i_level in reversed(range(n_levels)):
        for i_block in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, f"{p}.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = upsample(sd, f"{p}.up.{i_level}.upsample", h)
    h = nonlinearity(normalize(sd, p + ".norm_out", h))
    return F.conv2d(h, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def autoencoder_decode(sd, cfg: dict, z: Tensor) -> Tensor:
    """AutoencoderKL.decode (autoencoder.py:313-316): 1x1 post_quant_conv, then the decoder."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder(sd, cfg, z)


def decode_first_stage(sd, cfg: dict, z: Tensor, scale_factor: float = SDXL_SCALE_FACTOR) -> Tensor:
    """DiffusionEngine.decode_first_stage (diffusion.py:207-212): z / scale_factor, decode.  The
    caller maps to pixels with clamp((x + 1) / 2, 0, 1) (sample.py:195)."""
    return autoencoder_decode(sd, cfg, z * (1.0 / scale_factor))


def param_shapes(cfg: dict, embed_dim: int = 4) -> Dict[str, tuple]:
    """Names / shapes of the decode-side parameters of AutoencoderKL (autoencoder.py:296-299,
    model.py:605-701) — the `first_stage_model.*` keys of sdxl_vae.safetensors used by decode."""
    ch, mult, nres, zc = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"], cfg["z_channels"]
    out: Dict[str, tuple] = {}

    def conv(name, cin, cout, k):
        out[name + ".weight"] = (cout, cin, k, k)
        out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    conv("post_quant_conv", embed_dim, zc, 1)
    block_in = ch * mult[-1]
    conv("decoder.conv_in", zc, block_in, 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for i_level in reversed(range(len(mult))):
        block_out = ch * mult[i_level]
        for i_block in range(nres + 1):
            res(f"decoder.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            conv(f"decoder.up.{i_level}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, cfg["out_ch"], 3)
    return out


def synthetic_state_dict(cfg: dict, seed: int = 0, embed_dim: int = 4) -> Dict[str, Tensor]:
    """Seeded decode-side weights with activations of O(1) through the stack: convs ~ N(0, 1/fan_in),
    norm scales near 1, small biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg, embed_dim).items():
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif ".norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


def decode_flops(cfg: dict, latent: int, embed_dim: int = 4) -> float:
    """2*MAC of one decode of a [1, 4, latent, latent] latent (convs + the mid attention)."""
    shapes = param_shapes(cfg, embed_dim)
    n_levels = len(cfg["ch_mult"])
    res_of = {}
    for name in shapes:
        if not name.endswith(".weight") or len(shapes[name]) != 4:
            continue
        if ".up." in name:
            lvl = int(name.split(".up.")[1].split(".")[0])
            r = latent * 2 ** (n_levels - 1 - lvl)
            if ".upsample." in name:
                r *= 2
        elif "norm_out" in name or "conv_out" in name:
            r = latent * 2 ** (n_levels - 1)
        else:
            r = latent
        res_of[name] = r
    fl = 0.0
    for name, r in res_of.items():
        co, ci, k, _ = shapes[name]
        fl += 2.0 * r * r * co * ci * k * k
    c = cfg["ch"] * cfg["ch_mult"][-1]
    fl += 4.0 * (latent * latent) ** 2 * c
    return fl
