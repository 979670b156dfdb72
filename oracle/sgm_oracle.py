"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 restatement of the reference's pose-conditioned SDXL UNet denoising step
(customdiffusion360/custom-diffusion360 @ 1a23f97), written from the reference's algorithm, one
function per reference symbol with its file:line.  It is functional (state dict + config in,
tensors out), NCHW, fp32, unfused and unhoisted — i.e. it follows the reference's arithmetic
literally so that it is an independent check of the restructured CUDA path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product package (custom_diffusion360_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against the reference's OWN modules imported in place from /root/reference
(oracle/ref_harness.py, tests/test_oracle_vs_reference.py — runs where the reference checkout
exists) and against golden vectors generated from those modules and committed under
tests/golden/ (tests/golden/make_golden.py).  Third-party semantics that are not vendored in the
reference (pytorch3d cameras, xformers attention) are restated from their documented conventions
(SURVEY.md §8c): that part is "parity unpinned" beyond hand-computed known answers
(tests/test_oracle_cameras.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------------
# schedule / denoiser / guider / sampler
# --------------------------------------------------------------------------------------------


def make_beta_schedule_linear(n_timestep: int, linear_start: float, linear_end: float) -> np.ndarray:
    """sgm/modules/diffusionmodules/util.py:19-32 ("linear": linspace of sqrt(beta), squared), float64."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def legacy_ddpm_sigmas(n: int, num_timesteps: int = 1000, linear_start: float = 0.00085,
                       linear_end: float = 0.0120, do_append_zero: bool = True, flip: bool = False) -> Tensor:
    """LegacyDDPMDiscretization (discretizer.py:42-69) through Discretization.__call__ (:17-21)."""
    betas = make_beta_schedule_linear(num_timesteps, linear_start, linear_end)
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    if n < num_timesteps:
        # generate_roughly_equally_spaced_steps, discretizer.py:11-14
        timesteps = np.linspace(num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
        alphas_cumprod = alphas_cumprod[timesteps]
    elif n != num_timesteps:
        raise ValueError
    sigmas = torch.tensor((1 - alphas_cumprod) / alphas_cumprod, dtype=torch.float32) ** 0.5
    sigmas = torch.flip(sigmas, (0,))
    if do_append_zero:
        sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
    return sigmas if not flip else torch.flip(sigmas, (0,))


def append_dims(x: Tensor, target_dims: int) -> Tensor:
    """sgm/util.py:192-199"""
    return x[(...,) + (None,) * (target_dims - x.ndim)]


class DiscreteDenoiserOracle:
    """DiscreteDenoiser + EpsScaling (denoiser.py:22-79, denoiser_scaling.py:26-31), inference part."""

    def __init__(self, num_idx: int = 1000):
        # DiscreteDenoiser.__init__: discretization(num_idx, do_append_zero=False, flip=True)
        self.sigmas = legacy_ddpm_sigmas(num_idx, do_append_zero=False, flip=True)

    def sigma_to_idx(self, sigma: Tensor) -> Tensor:
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def __call__(self, network, inp: Tensor, sigma: Tensor, cond: dict, **kwargs):
        sigma = self.sigmas[self.sigma_to_idx(sigma)]  # possibly_quantize_sigma
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, inp.ndim)
        c_skip = torch.ones_like(sigma)
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = self.sigma_to_idx(sigma.clone().reshape(sigma_shape))  # quantize_c_noise
        predict, fg, alphas, rgbs = network(inp * c_in, c_noise, cond, **kwargs)
        return predict * c_out + inp * c_skip, fg, alphas, rgbs


def guider_prepare_inputs(x: Tensor, s: Tensor, c: dict, uc: dict, rows: int):
    """ScheduledCFGImgTextRef.prepare_inputs (guiders.py:116-133, rows=3) /
    VanillaCFGImgRef.prepare_inputs (:152-166, rows=2) for the inference case where the *_ref
    halves are empty (b == x.size(0))."""
    c_out = {}
    for k in c:
        b = uc[k].shape[0]
        uc1, uc2 = uc[k].split([x.size(0), b - x.size(0)])
        c1, c2 = c[k].split([x.size(0), b - x.size(0)])
        if rows == 3:
            c_out[k] = torch.cat((uc1, uc1, c1, uc2, c2, c2), 0)
        else:
            c_out[k] = torch.cat((uc1, c1, uc2, c2), 0)
    return torch.cat([x] * rows), torch.cat([s] * rows), c_out


def guider_combine(x: Tensor, rows: int, scale: float, scale_im: float = 0.0) -> Tensor:
    """guiders.py:111-114 / :147-150"""
    if rows == 3:
        x_u, x_ic, x_c = x.chunk(3)
        return x_u + scale * (x_c - x_ic) + scale_im * (x_ic - x_u)
    x_u, x_c = x.chunk(2)
    return x_u + scale * (x_c - x_u)


def euler_edm_sample(denoise_fn, x: Tensor, cond: dict, uc: dict, num_steps: int, rows: int = 3,
                     scale: float = 7.5, scale_im: float = 3.5, return_trajectory: bool = False):
    """EulerEDMSampler (sampling.py:44-58, 96-136, 314-318) with s_churn = 0 (gamma = 0).
    denoise_fn(x_in, sigma_in, c_in) -> denoised (first element of the denoiser tuple)."""
    sigmas = legacy_ddpm_sigmas(num_steps)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    traj = []
    for i in range(len(sigmas) - 1):
        sigma = s_in * sigmas[i]
        next_sigma = s_in * sigmas[i + 1]
        denoised = denoise_fn(*guider_prepare_inputs(x, sigma, cond, uc, rows))
        denoised = guider_combine(denoised, rows, scale, scale_im)
        d = (x - denoised) / append_dims(sigma, x.ndim)  # to_d, sampling_utils.py:39
        dt = append_dims(next_sigma - sigma, x.ndim)
        x = x + dt * d
        if return_trajectory:
            traj.append(x.clone())
    return (x, traj) if return_trajectory else x


# --------------------------------------------------------------------------------------------
# UNet building blocks
# --------------------------------------------------------------------------------------------


def timestep_embedding(timesteps: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """util.py:206-230"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm32(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """GroupNorm32 (util.py:309-311)"""
    return F.group_norm(x.float(), 32, w, b, eps)


def resblock(sd: Dict[str, Tensor], p: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward (openaimodel.py:350-376), no up/down, no scale-shift norm."""
    h = F.silu(group_norm32(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"]))
    h = F.conv2d(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + emb_out[:, :, None, None]
    h = F.silu(group_norm32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"]))
    h = F.conv2d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if (p + "skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def cross_attention(sd, p: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """MemoryEfficientCrossAttention.forward (attention.py:352-425); xformers' exact softmax
    attention restated with explicit matmuls (scale = dim_head ** -0.5)."""
    context = x if context is None else context
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(context, sd[p + "to_k.weight"])
    v = F.linear(context, sd[p + "to_v.weight"])
    b, nq, inner = q.shape
    dh = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, dh).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    out = torch.matmul(att, v).permute(0, 2, 1, 3).reshape(b, nq, inner)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def feed_forward(sd, p: str, x: Tensor) -> Tensor:
    """FeedForward with GEGLU (attention.py:89-115)"""
    h = F.linear(x, sd[p + "net.0.proj.weight"], sd[p + "net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), sd[p + "net.2.weight"], sd[p + "net.2.bias"])


def layer_norm(sd, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


# --------------------------------------------------------------------------------------------
# cameras (PyTorch3D PerspectiveCameras semantics, restated)
# packed camera row: R (9, row-major) | T (3) | focal (2) | principal point (2)
# --------------------------------------------------------------------------------------------


def pack_cameras(R: Tensor, T: Tensor, focal: Tensor, pp: Tensor) -> Tensor:
    return torch.cat([R.reshape(-1, 9), T.reshape(-1, 3), focal.reshape(-1, 2), pp.reshape(-1, 2)], -1).float()


def _cam_fields(cams: Tensor):
    return cams[..., :9].reshape(*cams.shape[:-1], 3, 3), cams[..., 9:12], cams[..., 12:14], cams[..., 14:16]


def camera_centers(cams: Tensor) -> Tensor:
    """get_camera_center: C = -T R^T (row-vector convention X_cam = X_world R + T)."""
    R, T, _, _ = _cam_fields(cams)
    return -torch.einsum("...j,...kj->...k", T, R)


def transform_points_ndc(cams: Tensor, pts: Tensor) -> Tensor:
    """cams [..., 16], pts [..., P, 3] -> NDC xy [..., P, 2]; no z clamp (inf/NaN behind camera)."""
    R, T, f, pp = _cam_fields(cams)
    v = torch.matmul(pts, R) + T[..., None, :]
    return f[..., None, :] * v[..., :2] / v[..., 2:3] + pp[..., None, :]


def unproject_ndc_depth1_dirs(cams: Tensor, xy: Tensor) -> Tensor:
    """Directions of get_directional_raybundle (utils_cameraray.py:61-100): unproject (x, y, 1),
    subtract the camera centre, normalise.  cams [..., 16], xy [P, 2] -> [..., P, 3]."""
    R, T, f, pp = _cam_fields(cams)
    vc = torch.cat([(xy - pp[..., None, :]) / f[..., None, :],
                    torch.ones(*pp.shape[:-1], xy.shape[0], 1)], -1)
    world = torch.matmul(vc - T[..., None, :], R.transpose(-1, -2))
    d = world - camera_centers(cams)[..., None, :]
    return d / d.norm(dim=-1, keepdim=True)


def patch_ray_xy(res: int) -> Tensor:
    """get_patch_raybundle, non-stratified (utils_cameraray.py:103-158): patch centres of
    linspace(1, -1, res+1), meshgrid indexing='xy', flattened row-major -> [res*res, 2]."""
    edges = torch.linspace(1, -1, res + 1)
    centers = (edges[:-1] + edges[1:]) / 2
    h_pos, v_pos = torch.meshgrid(centers, centers, indexing="xy")
    return torch.stack([h_pos.reshape(-1), v_pos.reshape(-1)], -1)


def raymarcher_depths(num_samples: int, far_plane: float, near_plane: float = 0.0):
    """Raymarcher.__init__/stratified_sampling, eval mode (nerfsd_pytorch3d.py:248-259, 326-330):
    bin centres of linspace(near, near+far, S+1) and bin widths.  NOTE NerfSDModule passes
    far_plane = near + far (:419)."""
    lengths = torch.linspace(near_plane, near_plane + far_plane, num_samples + 1)
    return (lengths[1:] + lengths[:-1]) / 2.0, lengths[1:] - lengths[:-1]


def positional_encoding(x: Tensor, n_freqs: int) -> Tensor:
    """utils_cameraray.py:222-242"""
    start = -1 * (n_freqs / 2)
    bands = 2.0 ** torch.arange(start, start + n_freqs) * np.pi
    return torch.cat([torch.sin(x * f) for f in bands] + [torch.cos(x * f) for f in bands], dim=-1)


def plucker(ray: Tensor) -> Tensor:
    """get_plucker_parameterization (utils_cameraray.py:201-219)"""
    o, d = ray[..., :3], ray[..., 3:]
    d = d / d.norm(dim=-1).unsqueeze(-1)
    return torch.cat([d, torch.cross(o, d, dim=-1)], dim=-1)


def stratified_ray_xy(res: int, rx: Tensor, ry: Tensor) -> Tensor:
    """get_patch_raybundle, stratified branch (utils_cameraray.py:111-140): each patch position is
    drawn uniformly between the centres of the neighbouring patches; rx / ry are the uniform
    variates [res+1] the reference draws with torch.rand_like (injected for reproducibility)."""
    def positions(r):
        edges = torch.linspace(1, -1, res + 1)
        center = (edges[1:] + edges[:-1]) / 2.0
        upper = torch.cat([center, edges[-1:]], -1)
        lower = torch.cat([edges[:1], center], -1)
        return (lower + (upper - lower) * r)[:-1]
    h_pos, v_pos = torch.meshgrid(positions(rx), positions(ry), indexing="xy")
    return torch.stack([h_pos.reshape(-1), v_pos.reshape(-1)], -1)


def stratified_depths(num_samples: int, far_plane: float, near_plane: float, t_rand: Tensor):
    """Raymarcher.stratified_sampling, training branch (nerfsd_pytorch3d.py:317-325): bin edges
    jittered between the neighbouring bin centres; t_rand [num_rays, S+1] uniform variates.
    Returns per-ray (depths [hw, S], dists [hw, S])."""
    lengths = torch.linspace(near_plane, near_plane + far_plane, num_samples + 1)
    center = (lengths[1:] + lengths[:-1]) / 2.0
    upper = torch.cat([center, lengths[-1:]], -1)
    lower = torch.cat([lengths[:1], center], -1)
    jit = lower[None] + (upper - lower)[None] * t_rand
    return (jit[:, :-1] + jit[:, 1:]) / 2.0, jit[:, 1:] - jit[:, :-1]


class _TruncExp(torch.autograd.Function):
    """attention.py:192-208: exp forward, gradient exp(clamp(x, max=15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(max=15))


def feature_nerf_encoding(sd, p: str, cams: Tensor, xref: Tensor, num_samples: int, far: float,
                          near: float = 0.0, num_freqs: int = 16, rgb_predict: bool = True,
                          jitter: Optional[dict] = None, mask_ref: Optional[Tensor] = None):
    """Raymarcher.forward (eval, prev_weights=None; nerfsd_pytorch3d.py:332-394) +
    FeatureNeRFEncoding.forward (:53-161) + the split in NerfSDModule.forward (:443-449).
    cams [b, n+1, 16] (index 0 = target), xref [b, n, hw, c].
    Returns features [b, hw, d, c], raw rgb [b, hw, d, 3] | None, raw sigma [b, hw, d, 1],
    dists [1, hw, d, 1], view softmax [b, n, hw, d, 1]."""
    b, n, hw, c = xref.shape
    res = int(math.sqrt(hw))
    d = num_samples
    if mask_ref is not None:
        # :61-70 — padding mask of every reference view [b, n, 1, H, W], nearest-resized to the block's
        # resolution, zeroes the reference tokens that lie in the padded border
        m = F.interpolate(mask_ref.reshape(b * n, *mask_ref.shape[2:]).float(), size=[res, res], mode="nearest")
        xref = xref * m.reshape(b, n, -1, 1)
    xy = patch_ray_xy(res) if jitter is None else stratified_ray_xy(res, *jitter["xy_rand"])
    centers = camera_centers(cams)                                   # [b, n+1, 3]
    dirs = unproject_ndc_depth1_dirs(cams, xy)                       # [b, n+1, hw, 3]
    rays = torch.cat([centers[:, :, None, :].expand(-1, -1, hw, -1), dirs], -1)  # [b, n+1, hw, 6]
    if jitter is None:
        depths, deltas = raymarcher_depths(d, near + far, near)
        depths, deltas = depths[None].expand(hw, d), deltas[None].expand(hw, d)
    else:  # training with stratified=True: per-ray jittered bins
        depths, deltas = stratified_depths(d, near + far, near, jitter["t_rand"])
    # ray_bundle_to_ray_points on the TARGET rays only (:381-387)
    ray_points = rays[:, 0, :, None, :3] + depths[None, :, :, None] * rays[:, 0, :, None, 3:]  # [b,hw,d,3]
    dists = deltas[None, :, :, None]

    R, T, _, _ = _cam_fields(cams)
    # :72-98 project into every camera, sample the reference feature maps
    ndc = transform_points_ndc(cams, ray_points.reshape(b, 1, hw * d, 3).expand(-1, n + 1, -1, -1))
    grid = torch.clip(torch.nan_to_num(-1 * ndc[:, 1:]), -1.2, 1.2).reshape(b * n, hw, d, 2)
    fmap = xref.reshape(b * n, res, res, c).permute(0, 3, 1, 2)
    plane = F.grid_sample(fmap, grid, align_corners=True, padding_mode="zeros")  # [bn, c, hw, d]
    plane = plane.reshape(b, n, c, hw, d).permute(0, 1, 3, 4, 2)                  # [b, n, hw, d, c]

    # :102-123 geometry features
    p_view = torch.einsum("bsdj,bnjk->bnsdk", ray_points, R) + T[:, :, None, None, :]  # [b,n+1,hw,d,3]
    p_view_enc = positional_encoding(p_view, num_freqs)
    o_t, d_t = rays[:, 0, :, :3], rays[:, 0, :, 3:]
    ray_view = torch.cat([torch.einsum("bsj,bnjk->bnsk", o_t, R) + T[:, :, None, :],
                          torch.einsum("bsj,bnjk->bnsk", d_t, R)], -1)[:, 1:]           # [b,n,hw,6]
    ray_view = ray_view[:, :, :, None, :].expand(-1, -1, -1, d, -1)
    ray_view_enc = positional_encoding(plucker(ray_view), num_freqs // 2)
    # convert_to_target_space(pose, rays[:, 1:])[..., :3]: reference origins in the target frame
    o_ref_t = torch.einsum("bnsj,bjk->bnsk", rays[:, 1:, :, :3], R[:, 0]) + T[:, 0, None, None, :]
    o_ref_t = o_ref_t[:, :, :, None, :].expand(-1, -1, -1, d, -1)
    o_ref_t_enc = positional_encoding(o_ref_t, num_freqs)

    mlp_in = torch.cat([plane, p_view_enc[:, 1:], p_view[:, 1:], ray_view_enc, ray_view[..., 3:]], -1)
    h = F.linear(mlp_in, sd[p + "plane_coefs.0.weight"], sd[p + "plane_coefs.0.bias"])
    h = F.linear(F.silu(h), sd[p + "plane_coefs.2.weight"], sd[p + "plane_coefs.2.bias"])
    nv_in = torch.cat([plane, p_view_enc[:, :1].expand(-1, n, -1, -1, -1),
                       p_view[:, :1].expand(-1, n, -1, -1, -1), o_ref_t, o_ref_t_enc], -1)
    attn = torch.softmax(F.linear(nv_in, sd[p + "nviews.weight"], sd[p + "nviews.bias"]), dim=1)
    final = (h * attn).sum(1)                                         # [b, hw, d, c]
    out = F.linear(final, sd[p + "decoder.weight"])                  # [b, hw, d, 4] (rgb3, sigma1)
    sigma_raw = out[..., -1:]
    rgb_raw = out[..., :-1] if rgb_predict else None
    return final, rgb_raw, sigma_raw, dists, attn


def vol_render(features: Tensor, densities: Tensor, dists: Tensor, rgb: Optional[Tensor]):
    """VolRender.get_weights/forward (nerfsd_pytorch3d.py:170-231)"""
    dd = dists * densities
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((*trans.shape[:2], 1, 1)), trans], dim=-2)
    weights = torch.nan_to_num(alphas * torch.exp(-trans))
    fg = torch.sum(weights, -2)
    rendered = torch.sum(weights * features, dim=-2)
    if rgb is not None:
        rgb = torch.sum(weights * rgb, dim=-2)
    return rendered, fg, alphas, rgb


def reference_attn(sd, p: str, cams: Tensor, xref: Tensor, context: Tensor, heads: int, cfg: dict):
    """BasicTransformerBlock.reference_attn (attention.py:571-598): FeatureNeRF, the block's own
    norm2/attn2 over every depth sample, trunc_exp, volume rendering, sigmoid on rgb."""
    jit = cfg.get("_jitter")            # training: iterator over the per-block injected variates
    feats, rgb_raw, sigma_raw, dists, _ = feature_nerf_encoding(
        sd, p + "pose_featurenerf.model.", cams, xref, cfg["num_samples"], cfg.get("far", 2.0),
        cfg.get("near_plane", 0.0), cfg.get("num_freqs", 16), cfg.get("rgb_predict", True),
        jitter=next(jit) if jit is not None else None, mask_ref=cfg.get("_mask_ref"))
    b, hw, d, c = feats.shape
    f2 = feats.reshape(b, hw * d, c)
    f2 = cross_attention(sd, p + "attn2.", layer_norm(sd, p + "norm2.", f2), context, heads) + f2
    feats = f2.reshape(b, hw, d, c)
    rendered, fg, alphas, rgb = vol_render(feats, _TruncExp.apply(sigma_raw), dists,
                                           torch.sigmoid(rgb_raw) if rgb_raw is not None else None)
    return rendered, fg, alphas, rgb


def build_context_ref(references: Tensor, choices: Sequence[int], batch: int) -> Tensor:
    """_customforward (sample.py:85-96): CFG row 0 sees the 'null' reference (last row of the
    stored buffer) repeated n times, the other rows see the chosen real references."""
    rows = 3 if batch % 3 == 0 else 2
    bs = batch // rows
    real = torch.stack([references[:-1][y] for y in choices]).unsqueeze(0).expand(bs, -1, -1, -1)
    null = references[-1:].unsqueeze(0).expand(bs, real.size(1), -1, -1)
    return torch.cat([null] + [real] * (rows - 1), dim=0)


def transformer_block(sd, p: str, x: Tensor, context: Tensor, heads: int, cfg: dict,
                      cams: Optional[Tensor] = None, choices=None, cache: Optional[dict] = None,
                      capture: Optional[dict] = None, ctx_ref: Optional[dict] = None):
    """BasicTransformerBlock as run at inference: _customforward (sample.py:82-136).
    capture: reference-stream mode — the block runs WITHOUT pose conditioning
    (attention.py:852-853: `xr = block(xr, context=contextr[i])`) and the output tokens of every
    block that owns pose weights are recorded under its prefix, which is what the validation hook
    stores as `references` (diffusion.py:28-40 keeps outputs whose fg_mask is None; main.py:594-602)."""
    aux = None
    x = cross_attention(sd, p + "attn1.", layer_norm(sd, p + "norm1.", x), None, heads) + x
    x = cross_attention(sd, p + "attn2.", layer_norm(sd, p + "norm2.", x), context, heads) + x
    if (p + "pose_emb_layers.weight") in sd and cams is not None:
        if cache is not None and p in cache:
            xref = cache[p]
        else:
            if ctx_ref is not None and p in ctx_ref:  # live reference stream (attention.py:852-854)
                cref = ctx_ref[p]
            else:
                cref = build_context_ref(sd[p + "references"], choices, x.size(0))
            xref, fg, alphas, rgb = reference_attn(sd, p, cams, cref, context, heads, cfg)
            aux = (fg, alphas, rgb)
            if cache is not None:
                cache[p] = xref
        cat_in = torch.cat([x, xref], -1)
        x = F.linear(cat_in, sd[p + "pose_emb_layers.weight"])
        probe = cfg.get("_probe")       # tests: (input, output) of pose_emb_layers, output keeps its grad
        if probe is not None:
            if x.requires_grad:
                x.retain_grad()
            probe[p] = (cat_in, x)
    x = feed_forward(sd, p + "ff.", layer_norm(sd, p + "norm3.", x)) + x
    if capture is not None and (p + "pose_emb_layers.weight") in sd:
        capture[p] = x
    return x, aux


def spatial_transformer(sd, p: str, x: Tensor, context: Tensor, heads: int, depth: int, cfg: dict,
                        image_cross: bool, cams=None, choices=None, cache=None, aux_out=None,
                        capture=None, ctx_ref=None):
    """SpatialTransformer at inference: customforward (sample.py:33-79), use_linear=True."""
    b, c, h, w = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)  # Normalize, attention.py:118
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = F.linear(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    interval = cfg.get("poscontrol_interval", 4)
    for i in range(depth):
        pose_here = image_cross and (i % interval == 0)
        x, aux = transformer_block(sd, f"{p}transformer_blocks.{i}.", x, context, heads, cfg,
                                   cams if pose_here else None, choices, cache, capture, ctx_ref)
        if aux is not None and aux_out is not None:
            aux_out.append(aux)
    x = F.linear(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


# --------------------------------------------------------------------------------------------
# UNet topology + forward
# --------------------------------------------------------------------------------------------


def unet_layout(cfg: dict) -> dict:
    """The module list UNetModel.__init__ builds (openaimodel.py:717-973), as plain tuples:
    ('conv_in',), ('res', cin, cout), ('st', ch, heads, depth, image_cross), ('down', ch),
    ('up', ch).  Attention ids are assigned in construction order (:728,791,877,944)."""
    mc = cfg["model_channels"]
    mult = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    nrb = [nrb] * len(mult) if isinstance(nrb, int) else list(nrb)
    attn_res = list(cfg["attention_resolutions"])
    depth = cfg["transformer_depth"]
    depth = [depth] * len(mult) if isinstance(depth, int) else list(depth)
    nhc = cfg["num_head_channels"]
    cross = list(cfg.get("image_cross_blocks") or [])
    inputs: List[list] = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds, aid = mc, 1, 0
    for level, m in enumerate(mult):
        for _ in range(nrb[level]):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("st", ch, ch // nhc, depth[level], aid in cross))
                aid += 1
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("st", ch, ch // nhc, depth[-1], aid in cross), ("res", ch, ch)]
    aid += 1
    outputs: List[list] = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb[level] + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in attn_res:
                layers.append(("st", ch, ch // nhc, depth[level], aid in cross))
                aid += 1
            if level and i == nrb[level]:
                layers.append(("up", ch))
                ds //= 2
            outputs.append(layers)
    return {"input_blocks": inputs, "middle_block": middle, "output_blocks": outputs, "out_ch": ch}


def _run_layers(sd, prefix, layers, h, emb, context, cfg, cams, choices, cache, aux_out, capture=None,
                ctx_ref=None):
    for j, layer in enumerate(layers):
        p = f"{prefix}{j}."
        kind = layer[0]
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        elif kind == "res":
            h = resblock(sd, p, h, emb)
        elif kind == "st":
            _, ch, heads, depth, image_cross = layer
            h = spatial_transformer(sd, p, h, context, heads, depth, cfg, image_cross, cams, choices,
                                    cache, aux_out, capture, ctx_ref)
        elif kind == "down":  # Downsample (openaimodel.py:183-230)
            h = F.conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)
        elif kind == "up":  # Upsample (openaimodel.py:114-164)
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
    return h


def unet_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, timesteps: Tensor, context: Tensor,
                 y: Tensor, cams: Optional[Tensor] = None, choices=None, cache: Optional[dict] = None,
                 capture: Optional[dict] = None, ctx_ref: Optional[dict] = None):
    """UNetModel.forward (openaimodel.py:975-1093) as executed by sample.py: no reference stream
    (input_ref absent), pose-enabled blocks take their reference tokens from the stored
    `references` buffers.  Returns (eps, aux) with aux = list of (fg_mask, alphas, rgb) per pose
    block that ran FeatureNeRF in this call.
    capture (dict, with cams=None): reference-stream pass over reference latents x — the tokens
    leaving every pose-capable block are stored as capture["<block prefix>"] = [B, hw, c]
    (the UNet's ref stream, openaimodel.py:79-111 / attention.py:830-868, shares all weights with
    the main stream and never sees pose conditioning)."""
    aux_out: list = []
    layout = unet_layout(cfg)
    t_emb = timestep_embedding(timesteps, cfg["model_channels"])
    emb = F.linear(F.silu(F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    emb = emb + F.linear(F.silu(F.linear(y, sd["label_emb.0.0.weight"], sd["label_emb.0.0.bias"])),
                         sd["label_emb.0.2.weight"], sd["label_emb.0.2.bias"])
    h = x
    hs = []
    for i, layers in enumerate(layout["input_blocks"]):
        h = _run_layers(sd, f"input_blocks.{i}.", layers, h, emb, context, cfg, cams, choices, cache, aux_out,
                        capture, ctx_ref)
        hs.append(h)
    h = _run_layers(sd, "middle_block.", layout["middle_block"], h, emb, context, cfg, cams, choices, cache, aux_out,
                    capture, ctx_ref)
    probe_h = cfg.get("_probe_h")       # tests: activations along the decoder, keeping their gradients
    def _keep(tag, t):
        if probe_h is not None and t.requires_grad:
            t.retain_grad()
            probe_h[tag] = t
    _keep("middle", h)
    for i, layers in enumerate(layout["output_blocks"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}.", layers, h, emb, context, cfg, cams, choices, cache, aux_out,
                        capture, ctx_ref)
        _keep(f"out{i}", h)
    h = F.silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1), aux_out


def unet_forward_with_reference_stream(sd, cfg: dict, x: Tensor, timesteps: Tensor, context: Tensor,
                                       y: Tensor, cams: Tensor, input_ref: Tensor, sigmas_ref: Tensor,
                                       context_ref: Tensor, y_ref: Tensor):
    """UNetModel.forward WITH `input_ref` (openaimodel.py:1008-1051, the training-time call shape,
    forward only): the reference latents input_ref [b, n, 4, L, L] run through the same weights as
    a second, pose-free stream whose time embedding is `timestep_embedding(sigmas_ref)` broadcast
    over the n views (:1041-1049) and whose text / vector conditioning are the second halves of
    context / y; every pose block of the main stream then reads the reference stream's tokens of
    that block as its context_ref (attention.py:852-854) instead of a stored `references` buffer.
    The reference stream carries no gradient: the reference runs it under torch.no_grad and hands
    `xr.detach()` to the pose blocks (openaimodel.py:95-108, attention.py:846-868).
    Returns ((eps, aux), captured reference tokens)."""
    b, n = input_ref.shape[:2]
    cap: dict = {}
    t_ref = sigmas_ref.reshape(b, 1).expand(b, n).reshape(b * n)
    with torch.no_grad():
        unet_forward(sd, cfg, input_ref.reshape(b * n, *input_ref.shape[2:]), t_ref, context_ref,
                     y_ref.reshape(b * n, -1), capture=cap)
    live = {p: v.detach().reshape(b, n, *v.shape[1:]) for p, v in cap.items()}
    return unet_forward(sd, cfg, x, timesteps, context, y, cams=cams, ctx_ref=live), cap


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (numpy RandomState: platform independent)
# --------------------------------------------------------------------------------------------

SDXL_CFG = dict(
    adm_in_channels=2816, num_classes="sequential", use_checkpoint=False, in_channels=4, out_channels=4,
    model_channels=320, attention_resolutions=[4, 2], num_res_blocks=2, channel_mult=[1, 2, 4],
    num_head_channels=64, use_linear_in_transformer=True, transformer_depth=[1, 2, 10], context_dim=2048,
    spatial_transformer_attn_type="softmax-xformers", image_cross_blocks=[0, 2, 4, 6, 8, 10], rgb=True,
    far=2, num_samples=24, not_add_context_in_triplane=False, rgb_predict=True, add_lora=False,
    average=False, use_prev_weights_imp_sample=True, stratified=True, imp_sampling_percent=0.9)

# same topology rules, 1/5 of the width: every code path of the SDXL config (3 levels, depth-2 and
# depth-3 transformers with pose blocks, skip concats, up/down) in a model that runs in seconds
TINY_CFG = dict(SDXL_CFG, adm_in_channels=96, model_channels=64, transformer_depth=[1, 2, 5],
                context_dim=128, image_cross_blocks=[0, 2, 4, 6, 8, 10], num_samples=6)


def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """Name -> shape of every parameter of the reference UNetModel for `cfg`, in reference
    state-dict naming (checked against the real module in tests/test_oracle_vs_reference.py)."""
    mc = cfg["model_channels"]
    ted = 4 * mc
    ctx = cfg["context_dim"]
    shapes: Dict[str, tuple] = {}

    def lin(p, o, i, bias=True):
        shapes[p + "weight"] = (o, i)
        if bias:
            shapes[p + "bias"] = (o,)

    def conv(p, o, i, k):
        shapes[p + "weight"] = (o, i, k, k)
        shapes[p + "bias"] = (o,)

    def norm(p, c):
        shapes[p + "weight"] = (c,)
        shapes[p + "bias"] = (c,)

    lin("time_embed.0.", ted, mc)
    lin("time_embed.2.", ted, ted)
    lin("label_emb.0.0.", ted, cfg["adm_in_channels"])
    lin("label_emb.0.2.", ted, ted)

    def add_layers(prefix, layers):
        for j, layer in enumerate(layers):
            p = f"{prefix}{j}."
            if layer[0] == "conv_in":
                conv(p, layer[2], layer[1], 3)
            elif layer[0] == "res":
                _, cin, cout = layer
                norm(p + "in_layers.0.", cin)
                conv(p + "in_layers.2.", cout, cin, 3)
                lin(p + "emb_layers.1.", cout, ted)
                norm(p + "out_layers.0.", cout)
                conv(p + "out_layers.3.", cout, cout, 3)
                if cin != cout:
                    conv(p + "skip_connection.", cout, cin, 1)
            elif layer[0] == "st":
                _, ch, heads, depth, image_cross = layer
                norm(p + "norm.", ch)
                lin(p + "proj_in.", ch, ch)
                for d in range(depth):
                    q = f"{p}transformer_blocks.{d}."
                    for a, cdim in (("attn1.", ch), ("attn2.", ctx)):
                        lin(q + a + "to_q.", ch, ch, bias=False)
                        lin(q + a + "to_k.", ch, cdim, bias=False)
                        lin(q + a + "to_v.", ch, cdim, bias=False)
                        lin(q + a + "to_out.0.", ch, ch)
                    lin(q + "ff.net.0.proj.", 8 * ch, ch)
                    lin(q + "ff.net.2.", ch, 4 * ch)
                    for nn_ in ("norm1.", "norm2.", "norm3."):
                        norm(q + nn_, ch)
                    if image_cross and d % cfg.get("poscontrol_interval", 4) == 0:
                        lin(q + "pose_emb_layers.", ch, 2 * ch, bias=False)
                        m = q + "pose_featurenerf.model."
                        lin(m + "plane_coefs.0.", ch, ch + 198)
                        lin(m + "plane_coefs.2.", ch, ch)
                        lin(m + "nviews.", 1, ch + 198)
                        lin(m + "decoder.", 4 if cfg.get("rgb_predict", True) else 1, ch, bias=False)
                lin(p + "proj_out.", ch, ch)
            elif layer[0] == "down":
                conv(p + "op.", layer[1], layer[1], 3)
            elif layer[0] == "up":
                conv(p + "conv.", layer[1], layer[1], 3)

    layout = unet_layout(cfg)
    for i, layers in enumerate(layout["input_blocks"]):
        add_layers(f"input_blocks.{i}.", layers)
    add_layers("middle_block.", layout["middle_block"])
    for i, layers in enumerate(layout["output_blocks"]):
        add_layers(f"output_blocks.{i}.", layers)
    norm("out.0.", layout["out_ch"])
    conv("out.2.", cfg["out_channels"], mc, 3)
    return shapes


def pose_block_prefixes(cfg: dict) -> List[tuple]:
    """(prefix, channels, downsample factor) of every FeatureNeRF block, in execution order."""
    out = []
    layout = unet_layout(cfg)

    def scan(prefix, layers, ds):
        for j, layer in enumerate(layers):
            if layer[0] == "st" and layer[4]:
                for d in range(layer[3]):
                    if d % cfg.get("poscontrol_interval", 4) == 0:
                        out.append((f"{prefix}{j}.transformer_blocks.{d}.", layer[1], ds))

    ds = 1
    for i, layers in enumerate(layout["input_blocks"]):
        if layers[0][0] == "down":
            ds *= 2
        scan(f"input_blocks.{i}.", layers, ds)
    scan("middle_block.", layout["middle_block"], ds)
    for i, layers in enumerate(layout["output_blocks"]):
        scan(f"output_blocks.{i}.", layers, ds)
        if layers[-1][0] == "up":
            ds //= 2
    return out


def synthetic_state_dict(cfg: dict, seed: int = 0, latent: Optional[int] = None,
                         num_references: int = 9) -> Dict[str, Tensor]:
    """Seeded weights for every parameter (numpy MT19937 -> identical on every platform).
    Linear/conv weights ~ N(0, 1/fan_in) so activations stay O(1) through the depth; norm gains
    ~ 1 + 0.1 N, biases 0.02 N.  Zero-initialised modules of the reference (proj_out, ResBlock
    out conv, `out`, decoder) are deliberately NON-zero here — with stock init the UNet output is
    exactly 0 and parity would be vacuous.  pose_emb_layers = [I | 0] + noise.
    With `latent` set, per-pose-block `references` buffers [num_references, hw, c] are added
    (sample.py:91-96 reads them; the last row is the 'null' reference)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("bias"):
            v = 0.02 * rs.standard_normal(shape)
        elif len(shape) == 1:
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rs.standard_normal(shape) / math.sqrt(fan_in)
            if name.endswith("pose_emb_layers.weight"):
                c = shape[0]
                v = 0.3 * v
                v[:, :c] += np.eye(c)
            if name.endswith("proj_out.weight") or name.endswith("out_layers.3.weight"):
                v = 0.5 * v  # residual branches: keep the residual stream from blowing up
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    if latent is not None:
        for prefix, c, ds in pose_block_prefixes(cfg):
            hw = (latent // ds) ** 2
            sd[prefix + "references"] = torch.from_numpy(
                rs.standard_normal((num_references, hw, c)).astype(np.float32))
    return sd


def lookat_cameras(n_views: int, seed: int = 0, radius: float = 1.5, focal: float = 2.0,
                   target_azimuth: float = 0.35) -> Tensor:
    """Target + n reference cameras on a circle looking at the origin, packed [n+1, 16]
    (PyTorch3D convention: R maps world->view by right-multiplication, +Z forward, +X left, +Y up)."""
    rs = np.random.RandomState(seed + 1234)
    az = [target_azimuth] + [2 * math.pi * k / n_views + 0.1 * rs.standard_normal() for k in range(n_views)]
    el = [0.25] + [0.2 + 0.1 * rs.standard_normal() for _ in range(n_views)]
    rows = []
    for a, e in zip(az, el):
        cpos = radius * np.array([math.cos(e) * math.sin(a), math.sin(e), math.cos(e) * math.cos(a)])
        z = -cpos / np.linalg.norm(cpos)                 # forward: towards the origin
        x = np.cross(np.array([0.0, 1.0, 0.0]), z)
        x /= np.linalg.norm(x)
        yv = np.cross(z, x)
        R = np.stack([x, yv, z], axis=1)                  # columns = view axes in world coords
        T = -cpos @ R
        rows.append(np.concatenate([R.reshape(-1), T, [focal, focal], [0.0, 0.0]]))
    return torch.from_numpy(np.stack(rows).astype(np.float32))


def synthetic_inputs(cfg: dict, latent: int, n_img: int = 1, seed: int = 0, n_views: int = 8) -> dict:
    """Seeded latents / text embeddings / vector cond / cameras for n_img images."""
    rs = np.random.RandomState(seed + 77)
    f = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    return dict(
        x=f(n_img, cfg["in_channels"], latent, latent),
        crossattn=f(n_img, 77, cfg["context_dim"]),
        vector=f(n_img, cfg["adm_in_channels"]),
        cams=torch.stack([lookat_cameras(n_views, seed + i, target_azimuth=0.35 + 0.5 * i) for i in range(n_img)]),
    )
