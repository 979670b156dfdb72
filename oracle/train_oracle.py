"""CPU oracle of the TRAINING step — TEST INFRASTRUCTURE ONLY (same rules as sgm_oracle.py).

Plain-PyTorch fp32 restatement of: StandardDiffusionLossImgRef.__call__/get_loss
(sgm/modules/diffusionmodules/loss.py:140-216), the training branch of Denoiser.__call__
(denoiser.py:22-44), CubicSampling / DiscreteSampling (sigma_sampling.py:16-53), the stratified
ray / depth jitter (utils_cameraray.py:111-140, nerfsd_pytorch3d.py:317-325, via cfg["_jitter"] in
sgm_oracle.reference_attn), _TruncExp (attention.py:192-208) and DiffusionEngine.forward's loss
assembly (sgm/models/diffusion.py:221-236).  Gradients come from torch.autograd over the functional
UNet of sgm_oracle (the state dict's pose tensors get requires_grad) — exactly what the reference
does with its modules.

Parity pinning: tests/test_oracle_vs_reference.py::test_training_step_vs_reference runs the
reference's OWN loss / denoiser / wrapper / UNetModel under autograd on the same weights and the
same random draws (replayed through torch's generator) and compares loss terms and every pose
gradient; tests/golden/train_step_golden.pt holds the vectors generated there.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import sgm_oracle as O

Tensor = torch.Tensor


def cubic_sigma_idx(u: Tensor, num_idx: int = 1000) -> Tensor:
    """CubicSampling.__call__ (sigma_sampling.py:47-53): idx = long((1 - u^3)(num_idx - 1))."""
    return ((1 - u ** 3) * (num_idx - 1)).long()


def training_sigmas(num_idx: int = 1000) -> Tensor:
    """The table both samplers index: discretization(num_idx, do_append_zero=False, flip=True)."""
    return O.legacy_ddpm_sigmas(num_idx, do_append_zero=False, flip=True)


def pose_param_names(sd: Dict[str, Tensor]):
    """trainkeys == 'pose' (diffusion.py:139-144): every parameter whose name contains 'pose'."""
    return [k for k in sd if "pose" in k and not k.endswith("references")]


def get_loss(model_output, fg_list, alphas_list, rgb_list, target, target_rgb, w, mask, opacity):
    """StandardDiffusionLossImgRef.get_loss, 'l2' (loss.py:173-206), incl. the cumulative re-resize
    of `opacity` inside the loop."""
    loss = w * (model_output - target) ** 2
    if mask is not None:
        loss_l2 = (loss * mask).sum([1, 2, 3]) / (mask.sum([1, 2, 3]) + 1e-6)
    else:
        loss_l2 = loss.reshape(target.shape[0], -1).mean(1)
    loss_fg, loss_bg, loss_rgb = [], [], []
    if len(fg_list) > 0 and len(alphas_list) > 0:
        for fg_mask, alphas in zip(fg_list, alphas_list):
            size = int(math.sqrt(fg_mask.size(1)))
            opacity = F.interpolate(opacity, size=size, antialias=True, mode="bilinear").detach()
            fg_mask = torch.clamp(fg_mask.reshape(-1, size * size), 0.0, 1.0)
            op = opacity.reshape(-1, size * size)
            loss_fg.append(((fg_mask - op) ** 2).mean(1))
            lb = (alphas - op[..., None, None]).abs() * (1 - op[..., None, None])
            loss_bg.append((lb * ((op[..., None, None] < 0.1) * 1)).mean([1, 2, 3]))
        loss_fg, loss_bg = torch.stack(loss_fg, 1), torch.stack(loss_bg, 1)
    if len(rgb_list) > 0 and target_rgb is not None:
        for rgb in rgb_list:
            size = int(math.sqrt(rgb.size(1)))
            mask_ = F.interpolate(mask, size=size, antialias=True, mode="bilinear").detach()
            tgt = F.interpolate(target_rgb * 0.5 + 0.5, size=size, antialias=True, mode="bilinear").detach()
            l = (tgt - rgb.reshape(-1, size, size, 3).permute(0, 3, 1, 2)) ** 2
            loss_rgb.append((l * mask_).sum([1, 2, 3]) / (mask.sum([1, 2, 3]) + 1e-6))
        loss_rgb = torch.stack(loss_rgb, 1)
    return loss_l2, loss_fg, loss_bg, loss_rgb


def training_loss(sd: Dict[str, Tensor], cfg: dict, batch: dict, *, global_step: int = 1,
                  lambdas=(10.0, 10.0, 5.0), rgb: bool = True, rgb_predict: bool = True):
    """One training-step loss.  batch: x [b,4,L,L], x_ref [b,n,4,L,L], cams [b,n+1,16], crossattn
    [b+b*n,77,ctx], vector [b+b*n,adm], mask [b,1,L,L] | None, opacity [b,1,H,W], rgb [b,3,H,W] |
    None, mask_ref [b,n,1,H,W] | None (padding masks of the reference views), drop_im [b], rand = {sigma_idx [b], sigma_ref_idx [b], noise, noise_ref, noise_ref2,
    jitter | None}.  Returns (total loss, dict of the terms)."""
    x, x_ref = batch["x"], batch["x_ref"]
    b, n = x_ref.shape[:2]
    rnd = batch["rand"]
    table = training_sigmas(1000)
    table_ref = training_sigmas(50)
    sigmas = table[rnd["sigma_idx"]]
    sigmas_ref = table_ref[rnd["sigma_ref_idx"]]
    noised = x + rnd["noise"] * O.append_dims(sigmas, x.ndim)
    # shared_step (diffusion.py:238-249): reference latents of samples with drop_im == 0 are zeroed
    x_ref = batch["drop_im"].reshape(b, 1, 1, 1, 1).float() * x_ref
    xr = x_ref + rnd["noise_ref"] * O.append_dims(sigmas_ref, x_ref.ndim)            # loss.py:163-170
    # Denoiser.__call__, training branch (denoiser.py:22-44); DiscreteDenoiser quantisation (:65-79)
    den = O.DiscreteDenoiserOracle(1000)
    sigma_q = den.sigmas[den.sigma_to_idx(sigmas)]
    xr = xr + rnd["noise_ref2"] * O.append_dims(sigmas_ref, xr.ndim)                 # :26-33 (second noising)
    c_in_ref = 1 / (O.append_dims(sigmas_ref, xr.ndim) ** 2 + 1.0) ** 0.5
    xr = xr * c_in_ref
    t_ref = den.sigma_to_idx(sigmas_ref)
    c_in = 1 / (sigma_q ** 2 + 1.0) ** 0.5
    c_noise = den.sigma_to_idx(sigma_q)
    jit = rnd.get("jitter")            # list of per-pose-block variates, in execution order
    cfg_run = dict(cfg, _jitter=iter(jit) if jit else None, _probe=batch.get("_probe"), _probe_h=batch.get("_probe_h"),
                   _mask_ref=batch.get("mask_ref"))          # loss.py:154 -> nerfsd_pytorch3d.py:61-70
    (eps, aux), _ = O.unet_forward_with_reference_stream(
        sd, cfg_run, noised * O.append_dims(c_in, x.ndim), c_noise, batch["crossattn"][:b], batch["vector"][:b],
        batch["cams"], xr, t_ref, batch["crossattn"][b:], batch["vector"][b:])
    model_output = eps * O.append_dims(-sigma_q, x.ndim) + noised
    w = O.append_dims(sigma_q ** -2.0, x.ndim)                                       # EpsWeighting
    fg_list = [a[0] for a in aux]
    al_list = [a[1] for a in aux]
    rgb_list = [a[2] for a in aux] if rgb_predict else []
    loss, loss_fg, loss_bg, loss_rgb = get_loss(model_output, fg_list, al_list, rgb_list, x, batch.get("rgb"), w,
                                                batch.get("mask"), batch["opacity"])
    # DiffusionEngine.forward (diffusion.py:221-236)
    drop = batch["drop_im"].reshape(-1).float()
    total = loss.mean()
    terms = {"loss": total.detach().clone()}
    lam_fg, lam_bg, lam_rgb = lambdas
    if rgb and global_step > 0:
        lf = (loss_fg.mean(1) * drop).sum() / (drop.sum() + 1e-12)
        lb = (loss_bg.mean(1) * drop).sum() / (drop.sum() + 1e-12)
        total = total + lam_fg * lf + lam_bg * lb
        terms.update(loss_fg=lf.detach(), loss_bg=lb.detach())
    if rgb_predict and torch.is_tensor(loss_rgb) and loss_rgb.mean() > 0:
        lr = (loss_rgb.mean(1) * drop).sum() / (drop.sum() + 1e-12)
        total = total + lam_rgb * lr
        terms["loss_rgb"] = lr.detach()
    return total, terms


def training_gradients(sd, cfg, batch, cond_grads: bool = False, **kw):
    """(total loss, terms, {pose parameter name: gradient}) by autograd over the functional oracle.
    cond_grads=True additionally returns, under the keys "cond.crossattn" / "cond.vector", the
    gradients w.r.t. the conditioner's outputs — what the reference's autograd hands back to the
    text encoders so that the `<new1>` token-embedding rows train (sgm/models/diffusion.py:343-356,
    main.py:627-643).  The rows of the reference views (b:) come out exactly zero: the reference
    stream runs under torch.no_grad (openaimodel.py:79-111, attention.py:846-868)."""
    names = pose_param_names(sd)
    sd = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    leaves = [sd[k] for k in names]
    if cond_grads:
        batch = dict(batch)
        batch["crossattn"] = batch["crossattn"].clone().requires_grad_(True)
        batch["vector"] = batch["vector"].clone().requires_grad_(True)
        leaves += [batch["crossattn"], batch["vector"]]
        names = names + ["cond.crossattn", "cond.vector"]
    total, terms = training_loss(sd, cfg, batch, **kw)
    grads = torch.autograd.grad(total, leaves, allow_unused=True)
    return total.detach(), terms, {k: (g if g is not None else torch.zeros_like(l))
                                   for k, g, l in zip(names, grads, leaves)}


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr=1e-4, betas=(0.9, 0.999), eps=1e-8,
               weight_decay=1e-2):
    """torch.optim.AdamW's update rule, restated (the reference's default optimizer_config)."""
    p = p * (1 - lr * weight_decay)
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    return p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps), m, v


def synthetic_train_batch(cfg: dict, latent: int, n_views: int = 4, b: int = 1, seed: int = 0, image: int = 64,
                          jitter: bool = False, mask_ref: bool = False) -> dict:
    """Seeded training batch of the shapes of config 4 (train_co3d_concept.yaml: 1 target + n refs)."""
    import numpy as np

    rs = np.random.RandomState(seed + 991)
    f = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    u = lambda *s: torch.from_numpy(rs.uniform(size=s).astype(np.float32))
    d = cfg["num_samples"]
    batch = dict(
        x=f(b, cfg["in_channels"], latent, latent), x_ref=f(b, n_views, cfg["in_channels"], latent, latent),
        cams=torch.stack([O.lookat_cameras(n_views, seed + i, target_azimuth=0.35 + 0.5 * i) for i in range(b)]),
        crossattn=f(b + b * n_views, 77, cfg["context_dim"]), vector=f(b + b * n_views, cfg["adm_in_channels"]),
        mask=(u(b, 1, latent, latent) > 0.25).float(), opacity=(u(b, 1, image, image) > 0.5).float(),
        rgb=u(b, 3, image, image) * 2 - 1, drop_im=torch.ones(b),
        rand=dict(sigma_idx=cubic_sigma_idx(u(b)), sigma_ref_idx=torch.from_numpy(rs.randint(0, 50, size=(b,))),
                  noise=f(b, cfg["in_channels"], latent, latent),
                  noise_ref=f(b, n_views, cfg["in_channels"], latent, latent),
                  noise_ref2=f(b, n_views, cfg["in_channels"], latent, latent)))
    if mask_ref:  # data_co3d.py:485 `masks_padding[1:]`: non-square images padded to a square -> border bands of 0
        m = torch.ones(b, n_views, 1, image, image)
        for i in range(b):
            for v in range(n_views):
                w = int(rs.randint(image // 16, image // 4))
                if (i + v) % 2:
                    m[i, v, :, :w, :] = 0; m[i, v, :, image - w:, :] = 0
                else:
                    m[i, v, :, :, :w] = 0; m[i, v, :, :, image - w:] = 0
        batch["mask_ref"] = m
    if jitter:   # one independent draw per pose block, in execution order (each Raymarcher call draws anew)
        jit = []
        for _, c, ds in O.pose_block_prefixes(cfg):
            res = latent // ds
            jit.append(dict(xy_rand=(u(res + 1), u(res + 1)), t_rand=u(res * res, d + 1)))
        batch["rand"]["jitter"] = jit
    return batch
