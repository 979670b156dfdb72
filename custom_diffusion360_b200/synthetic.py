"""Synthetic weights / inputs of the right shapes for benchmarks and smoke runs (there is no
network for checkpoints or datasets).  Everything is generated on the device.

SDXL_CFG is `network_config.params` of the reference's configs/train_co3d_concept.yaml:29-54.
"""
from __future__ import annotations

import math

import torch

SDXL_CFG = dict(
    adm_in_channels=2816, num_classes="sequential", use_checkpoint=False, in_channels=4, out_channels=4,
    model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4],
    num_head_channels=64, use_linear_in_transformer=True, transformer_depth=[1, 2, 10], context_dim=2048,
    spatial_transformer_attn_type="softmax-xformers", image_cross_blocks=[0, 2, 4, 6, 8, 10], rgb=True,
    far=2, num_samples=24, not_add_context_in_triplane=False, rgb_predict=True, add_lora=False,
    average=False, use_prev_weights_imp_sample=True, stratified=True, imp_sampling_percent=0.9)

# algorithmic FLOPs (2*MAC) of one UNet forward per batch row, FeatureNeRF excluded, counted on the
# reference model with torch.utils.flop_counter (SURVEY.md §8d / BASELINE.md §2)
UNET_TFLOP_PER_ROW = {64: 1.589, 128: 6.761}


@torch.no_grad()
def init_random_weights_(model: torch.nn.Module, seed: int = 0) -> None:
    """SDXL-shaped random weights: W ~ N(0, 1/fan_in) (residual-branch outputs halved), norm gains
    1 + 0.1 N, biases 0.02 N, pose_emb_layers = [I | 0] + 0.3 N/sqrt(fan_in).  The reference's
    zero-initialised modules are made non-zero on purpose (otherwise every residual branch is the
    identity and the network output is exactly zero)."""
    g = torch.Generator(device=next(model.parameters()).device)
    g.manual_seed(seed)
    for name, p in model.named_parameters():
        if name.endswith("bias"):
            p.normal_(0.0, 0.02, generator=g)
        elif p.dim() == 1:
            p.normal_(1.0, 0.1, generator=g)
        else:
            fan_in = p[0].numel()
            std = 1.0 / math.sqrt(fan_in)
            if name.endswith("proj_out.weight") or name.endswith("out_layers.3.weight"):
                std *= 0.5
            if name.endswith("pose_emb_layers.weight"):
                p.normal_(0.0, 0.3 * std, generator=g)
                c = p.shape[0]
                p[:, :c] += torch.eye(c, device=p.device, dtype=p.dtype)
            else:
                p.normal_(0.0, std, generator=g)


def lookat_cameras(n_views: int, seed: int = 0, radius: float = 1.5, focal: float = 2.0,
                   target_azimuth: float = 0.35) -> torch.Tensor:
    """Target + n reference cameras on a circle looking at the origin, packed fp32 [n+1, 16]
    (R row-major 9 | T 3 | focal 2 | principal point 2; PyTorch3D convention X_cam = X_world R + T)."""
    g = torch.Generator().manual_seed(seed + 1234)
    az = [target_azimuth] + [2 * math.pi * k / n_views + 0.1 * float(torch.randn((), generator=g))
                             for k in range(n_views)]
    el = [0.25] + [0.2 + 0.1 * float(torch.randn((), generator=g)) for _ in range(n_views)]
    rows = []
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    for a, e in zip(az, el):
        cpos = radius * torch.tensor([math.cos(e) * math.sin(a), math.sin(e), math.cos(e) * math.cos(a)],
                                     dtype=torch.float64)
        z = -cpos / cpos.norm()
        x = torch.linalg.cross(up, z)
        x = x / x.norm()
        y = torch.linalg.cross(z, x)
        R = torch.stack([x, y, z], dim=1)
        T = -cpos @ R
        rows.append(torch.cat([R.reshape(-1), T, torch.tensor([focal, focal, 0.0, 0.0], dtype=torch.float64)]))
    return torch.stack(rows).float()


def make_conditioning(cfg: dict, n_img: int, device, seed: int = 0):
    """(cond, uc): text-embedding-shaped tensors; uc uses zero text embeddings like
    `force_uc_zero_embeddings` (sample.py:155-161)."""
    g = torch.Generator(device=device).manual_seed(seed + 77)
    ca = torch.randn(n_img, 77, cfg["context_dim"], device=device, generator=g)
    vec = torch.randn(n_img, cfg["adm_in_channels"], device=device, generator=g)
    uvec = vec.clone()
    uvec[:, : cfg["adm_in_channels"] // 2] = 0
    return {"crossattn": ca, "vector": vec}, {"crossattn": torch.zeros_like(ca), "vector": uvec}


def make_references(model, latent: int, n_refs: int, device, seed: int = 0) -> dict:
    """Per-pose-block `references` buffers [n_refs + 1, hw, c] (last row = 'null' reference)."""
    g = torch.Generator(device=device).manual_seed(seed + 99)
    refs = {}
    # the token count of each pose block follows from its channel width (level) in the SDXL layout
    mc = model.model_channels
    for name, block in model.pose_blocks():
        c = block.pose_emb_layers.weight.shape[0]
        ds = c // mc  # 640 -> /2, 1280 -> /4
        hw = (latent // ds) ** 2
        refs[name] = torch.randn(n_refs + 1, hw, c, device=device, generator=g)
    return refs


# ---- first stage (VAE decode) ---------------------------------------------------
# `first_stage_config.params.ddconfig` / `scale_factor` of configs/train_co3d_concept.yaml:5,104-115
SDXL_VAE_DDCONFIG = dict(attn_type="vanilla-xformers", double_z=True, z_channels=4, resolution=256, in_channels=3,
                         out_ch=3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
SDXL_SCALE_FACTOR = 0.13025


@torch.no_grad()
def init_random_vae_weights_(vae: torch.nn.Module, seed: int = 0) -> None:
    """Decode-side weights with O(1) activations through the stack: convs ~ N(0, 1/fan_in), norm
    gains 1 + 0.1 N, biases 0.05 N."""
    g = torch.Generator(device=next(vae.parameters()).device)
    g.manual_seed(seed)
    for name, p in vae.named_parameters():
        if p.dim() == 4:
            p.normal_(0.0, 1.0 / math.sqrt(p[0].numel()), generator=g)
        elif name.endswith("weight"):
            p.normal_(1.0, 0.1, generator=g)
        else:
            p.normal_(0.0, 0.05, generator=g)


def vae_decode_flops(vae, latent: int) -> float:
    """2*MAC of one decode of a [1, 4, latent, latent] latent: every convolution at the resolution it
    runs at + the single-head mid attention (4 * hw^2 * C)."""
    dec = vae.decoder
    n_levels = dec.num_resolutions
    fl = 0.0
    for name, m in dec.named_modules():
        if not isinstance(m, torch.nn.Conv2d):
            continue
        r = latent
        if name.startswith("up."):
            lvl = int(name.split(".")[1])
            r = latent * 2 ** (n_levels - 1 - lvl) * (2 if ".upsample." in name else 1)
        elif name.startswith("conv_out"):
            r = latent * 2 ** (n_levels - 1)
        fl += 2.0 * r * r * m.out_channels * m.in_channels * m.kernel_size[0] * m.kernel_size[1]
    c = dec.mid.attn_1.in_channels
    return fl + 4.0 * float(latent * latent) ** 2 * c
