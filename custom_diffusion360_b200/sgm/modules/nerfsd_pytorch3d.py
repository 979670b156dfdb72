"""FeatureNeRF module surface (NerfSDModule / Raymarcher / FeatureNeRFEncoding / VolRender) on the
sm_100a kernels.  Class names, constructor kwargs, registry (`MODES`) and state-dict keys follow the
reference (sgm/modules/nerfsd_pytorch3d.py:23-464); the arithmetic is restructured as described in
csrc/nerf.cu (first Linear hoisted through the bilinear gather, second Linear applied after the
view-weighted sum) and runs on tensor cores + fused geometry kernels.

Scope: the flow the reference actually executes — deterministic depth bins at inference, injected
stratified variates in training (train_path.nerf_bins), no importance sampling (`prev_weights` is
never forwarded to the raymarcher); the reference padding masks `mask_ref`
(None at inference, sample.py:181; always present in training, data_co3d.py:485) are applied by
`apply_mask_ref` (nerfsd_pytorch3d.py:61-70).  Requests outside it raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from .utils_cameraray import pack_pose, patch_ray_xy

KPE = 208  # 198 positional features padded to a multiple of 8 (16-byte rows)


class Raymarcher(nn.Module):
    """Depth bins of the target rays (reference :234-330).  Buffers are registered under the
    reference's names so base checkpoints load with identical keys."""

    def __init__(self, num_samples=32, far_plane=2.0, stratified=False, training=True,
                 imp_sampling_percent=0.9, near_plane=0.0):
        super().__init__()
        self.num_samples = num_samples
        self.far_plane = far_plane
        self.near_plane = near_plane
        u = torch.linspace(0, 1 - 1.0 / num_samples, num_samples)
        lengths = torch.linspace(near_plane, near_plane + far_plane, num_samples + 1)
        center = (lengths[1:] + lengths[:-1]) / 2.0
        self.register_buffer("u", u)
        self.register_buffer("lengths", lengths)
        self.register_buffer("lengths_center", center)
        self.register_buffer("lengths_upper", torch.cat([center, lengths[-1:]], -1))
        self.register_buffer("lengths_lower", torch.cat([lengths[:1], center], -1))
        self.stratified = stratified
        self.imp_sampling_percent = imp_sampling_percent

    def bins(self, num_rays: int, device):
        """(depths [hw, d], dists [hw, d]) — stratified_sampling's deterministic branch (:326-330)."""
        lengths = self.lengths.to(device=device, dtype=torch.float32)
        depths = ((lengths[1:] + lengths[:-1]) / 2.0)[None].expand(num_rays, -1).contiguous()
        dists = (lengths[1:] - lengths[:-1])[None].expand(num_rays, -1).contiguous()
        return depths, dists


class FeatureNeRFEncoding(nn.Module):
    """Parameters of the per-sample MLP (reference :23-51): plane_coefs.{0,2}, nviews, decoder."""

    def __init__(self, in_channels, out_channels, far_plane: float = 2.0, rgb_predict=False,
                 average=False, num_freqs=16) -> None:
        super().__init__()
        if average:
            raise NotImplementedError("average=True (mean over views) is not built; the shipped config uses average: False")
        if num_freqs != 16:
            raise NotImplementedError("kernels are specialised for num_freqs=16 (198 positional features)")
        self.far_plane = far_plane
        self.rgb_predict = rgb_predict
        self.average = average
        self.num_freqs = num_freqs
        kin = in_channels + num_freqs * 3 * 4 + 2 * 3
        self.plane_coefs = nn.Sequential(nn.Linear(kin, out_channels), nn.SiLU(),
                                         nn.Linear(out_channels, out_channels))
        self.nviews = nn.Linear(kin, 1)
        self.decoder = nn.Linear(out_channels, 1 + (3 if rgb_predict else 0), bias=False)
        nn.init.zeros_(self.decoder.weight)  # zero_module, reference :49-51
        self._packed = None
        self._stale = True

    def packed(self):
        """bf16 operand packs of the (trainable) MLP; b1 / b2 / wnv_geo / bnv are live fp32 views of the
        parameters.  After `mark_stale()` the bf16 copies are refreshed IN PLACE (fixed addresses:
        CUDA-graph replays and the eager step share them, see attention._Packed)."""
        dev = self.nviews.weight.device
        p = self._packed
        if p is None or p["dev"] != dev:
            c = self.plane_coefs[2].weight.shape[0]
            p = dict(dev=dev, c=c, wg=torch.zeros(c + 8, c, device=dev, dtype=torch.bfloat16),
                     w1p=torch.zeros(c, KPE, device=dev, dtype=torch.bfloat16),
                     w2=torch.empty(c, c, device=dev, dtype=torch.bfloat16),
                     wd=torch.empty(self.decoder.weight.shape, device=dev, dtype=torch.bfloat16))
            self._packed = p
            self._stale = True
        if self._stale:
            c = p["c"]
            w1 = self.plane_coefs[0].weight.detach()
            wnv = self.nviews.weight.detach()
            # G projection: rows [0,c) = feature columns of plane_coefs.0, row c = feature columns of
            # nviews, zero rows up to c+8
            p["wg"][:c].copy_(w1[:, :c])
            p["wg"][c].copy_(wnv[0, :c])
            p["w1p"][:, :198].copy_(w1[:, c:])
            p["w2"].copy_(self.plane_coefs[2].weight.detach())
            p["wd"].copy_(self.decoder.weight.detach())
            p.update(b1=self.plane_coefs[0].bias.detach().float().contiguous(),
                     b2=self.plane_coefs[2].bias.detach().float().contiguous(),
                     wnv_geo=wnv[0, c:].detach().float().contiguous(),
                     bnv=self.nviews.bias.detach().float().contiguous())   # read on the device: no host sync
            self._stale = False
        return p

    def mark_stale(self):
        self._stale = True


class VolRender(nn.Module):
    """Stateless (reference :164-231); the scan runs in cd360_nerf_volrender."""

    def forward(self, features, densities, dists=None, rgb=None, **_):
        b, hw, d, c = features.shape
        feats = features.reshape(b * hw * d, c)
        feats = feats if feats.dtype == torch.bfloat16 else ops.cast_bf16(feats.float().contiguous())
        raw = torch.empty(b, hw, d, 4, device=features.device, dtype=torch.float32)
        # the kernel takes raw (pre-exp / pre-sigmoid) values: invert the activations the
        # reference applies outside VolRender (attention.py:590-594)
        raw[..., 3] = torch.log(densities[..., 0].float())
        raw[..., :3] = torch.logit(rgb.float()) if rgb is not None else 0.0
        dd = dists.reshape(-1, d)[:hw].float().contiguous()
        rendered, fg, alphas, rgb_o = ops.nerf_volrender(feats.contiguous(), raw, dd, b, hw, d, c)
        return (rendered.float().view(b, hw, c), fg.view(b, hw, 1), alphas.view(b, hw, d, 1), None,
                rgb_o if rgb is not None else None)


class NerfSDModule(nn.Module):
    MODES = {"feature-nerf": FeatureNeRFEncoding}

    def __init__(self, mode="feature-nerf", out_channels=None, far_plane=2.0, num_samples=32,
                 rgb_predict=False, average=False, num_freqs=16, stratified=False,
                 imp_sampling_percent=0.9, near_plane=0.0):
        super().__init__()
        self.rgb_predict = rgb_predict
        self.raymarcher = Raymarcher(num_samples=num_samples, far_plane=near_plane + far_plane,
                                     stratified=stratified, imp_sampling_percent=imp_sampling_percent,
                                     near_plane=near_plane)
        self.model = self.MODES[mode](out_channels, out_channels, far_plane=near_plane + far_plane,
                                      rgb_predict=rgb_predict, average=average, num_freqs=num_freqs)

    @staticmethod
    def apply_mask_ref(xref_tok: torch.Tensor, mask_ref, b: int, n: int, hw: int) -> torch.Tensor:
        """FeatureNeRFEncoding.forward step 1 (reference :61-70): `xref * nearest_resize(mask_ref)`.
        mask_ref [b, n, 1, H, W] (or [b*n, 1, H, W] / [b*n, H, W]); xref_tok bf16 [b*n*hw, c] -> new tensor."""
        if mask_ref is None:
            return xref_tok
        res = int(math.sqrt(hw))
        m = mask_ref.reshape(b * n, mask_ref.shape[-2], mask_ref.shape[-1])
        m = m.to(device=xref_tok.device, dtype=torch.float32).contiguous()
        return ops.nerf_mask_ref(xref_tok, m, b * n, res)

    # ---- token-layout fast path -------------------------------------------------------------
    def encode_tokens(self, cams: torch.Tensor, xref_tok: torch.Tensor, b: int, n: int, hw: int, mask_ref=None):
        """cams fp32 [b, n+1, 16]; xref_tok bf16 [b*n*hw, c] -> plane_features_final bf16
        [b*hw*d, c], raw fp32 [b*hw*d, 4|1], dists [hw, d], view softmax fp32 [b, n, hw*d]."""
        if self.training and self.raymarcher.stratified:
            raise NotImplementedError("stratified jitter needs injected variates: the training step goes through "
                                      "train_path.nerf_forward (UNetModel.forward_train)")
        xref_tok = self.apply_mask_ref(xref_tok, mask_ref, b, n, hw)
        pk = self.model.packed()
        c = pk["c"]
        d = self.raymarcher.num_samples
        res = int(math.sqrt(hw))
        assert res * res == hw
        dev = xref_tok.device
        xy = patch_ray_xy(res, dev)
        depths, dists = self.raymarcher.bins(hw, dev)
        g = ops.gemm(xref_tok, pk["wg"])                                   # [b*n*hw, c+8]
        pe, gidx, gwgt, vlogit = ops.nerf_points(cams, xy, depths, pk["wnv_geo"], pk["bnv"], b, n,
                                                 res, d, KPE)
        hpre = ops.gemm(pe, pk["w1p"], bias=pk["b1"])                      # [b*n*hw*d, c]
        del pe
        s, vsm = ops.nerf_combine(g, hpre, gidx, gwgt, vlogit, b, n, hw, d, c)
        del hpre, g
        final = ops.gemm(s, pk["w2"], bias=pk["b2"])                       # [b*hw*d, c]
        raw = ops.gemm(final, pk["wd"], out_fp32=True)                     # [b*hw*d, 4] (rgb3, sigma1)
        return final, raw, dists, vsm

    # ---- reference-signature entry point ----------------------------------------------------
    def forward(self, pose, xref=None, mask_ref=None, prev_weights=None, imp_sample_next_step=False):
        """Reference contract (:434-464): xref [b, n, hw, c] -> (features [b,hw,d,c], sigma_raw
        [b,hw,d,1], dists [1,hw,d,1], view softmax [b,n,hw,d,1], rgb_raw | None, None, None)."""
        b, n, hw, c = xref.shape
        cams = pack_pose(pose, xref.device)
        tok = xref.reshape(b * n * hw, c)
        tok = tok if tok.dtype == torch.bfloat16 else ops.cast_bf16(tok.float().contiguous())
        final, raw, dists, vsm = self.encode_tokens(cams, tok.contiguous(), b, n, hw, mask_ref=mask_ref)
        d = self.raymarcher.num_samples
        feats = ops.cast_f32(final).view(b, hw, d, c)
        raw = raw.view(b, hw, d, -1)
        rgb = raw[..., :3] if self.rgb_predict else None
        return (feats, raw[..., -1:], dists.view(1, hw, d, 1), vsm.view(b, n, hw, d, 1), rgb, None, None)
