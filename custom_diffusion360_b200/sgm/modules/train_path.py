"""Training forward + explicit backward of the pose-conditioned UNet on the sm_100a kernels.

Reference: `DiffusionEngine.training_step` -> `StandardDiffusionLossImgRef.__call__` -> denoiser ->
`UNetModel.forward` with `input_ref` (sgm/models/diffusion.py:221-272, loss.py:140-216,
openaimodel.py:975-1093), differentiated by torch.autograd.  Only the pose weights train
(`trainkeys: pose`, diffusion.py:139-144): `pose_emb_layers.weight` and the FeatureNeRF MLP
(`pose_featurenerf.model.{plane_coefs.0,plane_coefs.2,nviews,decoder}`) of every pose block.

There is no autograd here.  The main stream runs once in "taped" form (out-of-place residual
updates, the activations the backward needs are kept), then `unet_backward` walks the tape in
reverse and launches, per layer, the gradient kernels:
  * dX of every Linear / 1x1 / 3x3 conv = the tcgen05 GEMM over transposed / tap-flipped weight
    packs built once (`bwd_pack`);
  * attention, LayerNorm, GroupNorm+SiLU, GEGLU, resampling, FeatureNeRF gather / view-softmax /
    volume rendering: the dedicated backward kernels of csrc/train.cu and csrc/attention_bwd.cu;
  * dW of the (few, small) trainable Linears = GEMMs over transposed operands, fp32 out.
The reference-image stream is `no_grad` in the reference (attention.py:851-868) and stays a plain
forward here (`UNetModel.capture_references`).

Conditioning gradients (`cond_grad=True`, the default of the engine): the shipped config also trains
the `<new1>` rows of both text encoders' token tables (sgm/models/diffusion.py:343-356,
main.py:627-643), i.e. the reference's autograd carries dL/d(crossattn) and dL/d(vector) back into
the conditioner.  The walk then (a) also asks the attention backward of every attn2 — the 70
transformer blocks and the 12 `reference_attn` calls that reuse them — for dK / dV, collected in
one [b*77, sum(2*inner)] buffer that mirrors the fused K|V projection and turned into dL/d(context)
by ONE GEMM against the transposed K|V weights; (b) sums, per ResBlock, the gradient of
`emb_layers` over the pixels and walks it back through emb_layers / label_emb (SiLU backward
kernel + the small-linear kernel over transposed packs) to dL/d(vector); (c) does not stop at the
first pose block (every ResBlock upstream still feeds (b)).  Without it the walk stops at the first
pose block in forward order: nothing upstream of it is trainable.
"""
from __future__ import annotations

import contextlib
import os
import math
from types import SimpleNamespace as NS
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ... import ops
from ..prepack import pack_conv3x3_bwd, transposed
from .attention import BasicTransformerBlock, SpatialTransformer
from .nerfsd_pytorch3d import KPE
from .utils_cameraray import patch_ray_xy

bf16 = torch.bfloat16
f32 = torch.float32
# side stream of the current backward walk (unet_backward) and the tensors it still reads
_BWD: dict = {"side": None, "keep": []}
# FeatureNeRF results of the current taped forward that were enqueued ahead on their own stream:
# id(block) -> (rendered, aux, saved, event)
_FWD: dict = {}


# ------------------------------------------------------------------------------------------------
# backward weight packs (built lazily, once per module; invalidated with the forward packs)
# ------------------------------------------------------------------------------------------------
def bwd_pack(m) -> dict:
    p = m.__dict__.get("_bwdpk")
    dev = next(m.parameters()).device
    if p is not None and p["dev"] == dev:
        return p
    from .diffusionmodules import openaimodel as U  # late: avoid an import cycle

    p = dict(dev=dev)
    if isinstance(m, BasicTransformerBlock):
        a1, a2 = m.attn1, m.attn2
        p1, p2 = a1.packed(), a2.packed()
        p.update(wqkv_t=transposed(p1["wqkv"]), wo1_t=transposed(p1["wo"]),
                 wq2_t=transposed(p2["wq"]), wo2_t=transposed(p2["wo"]),
                 wff_t=transposed(m.ff.net[0].packed()["w"]),        # packed (interleaved) row order
                 w2_t=transposed(m.ff.net[2].packed()["w"]),
                 g1=m.norm1.packed()["g"], g2=m.norm2.packed()["g"], g3=m.norm3.packed()["g"])
    elif isinstance(m, SpatialTransformer):
        p.update(win_t=transposed(m.proj_in.packed()["w"]), wout_t=transposed(m.proj_out.packed()["w"]))
    elif isinstance(m, U.ResBlock):
        p.update(w1=pack_conv3x3_bwd(m.in_layers[2].weight.detach()),
                 w2=pack_conv3x3_bwd(m.out_layers[3].weight.detach()))
        fp = m.packed()
        if "ws" in fp:
            p["ws_t"] = transposed(fp["ws"])
    elif isinstance(m, U.Downsample):
        p["w_t"] = transposed(m.packed()["w"])
    elif isinstance(m, U.Upsample):
        p["w"] = pack_conv3x3_bwd(m.conv.weight.detach())
    elif isinstance(m, U.UNetModel):
        p["cout"] = pack_conv3x3_bwd(m.out[2].weight.detach(), cout_pad=64)
        if m.__dict__.get("_cond_grad"):   # conditioning gradients: dX packs of the K|V and embedding projections
            fp = m.packed()
            p.update(kvw_t=transposed(fp["kvw"]), embw_t=transposed(fp["embw"]), le2w_t=transposed(fp["le2w"]),
                     le0w_t=transposed(fp["le0w"]))
    else:
        raise TypeError(type(m))
    m.__dict__["_bwdpk"] = p
    return p


def pose_bwd_pack(block: BasicTransformerBlock) -> dict:
    """Transposed packs of the TRAINABLE weights of a pose block (rebuilt after every optimiser
    step, unlike the frozen packs above)."""
    p = block.__dict__.get("_bwdpk_pose")
    dev = block.pose_emb_layers.weight.device
    c = block.pose_emb_layers.weight.shape[0]
    if p is not None and p["dev"] == dev and not p.get("stale"):
        return p
    wp = block.pose_emb_layers.packed()["w"]
    pk = block.pose_featurenerf.model.packed()
    if p is None or p["dev"] != dev:
        p = dict(dev=dev, wp_x_t=torch.empty(c, c, device=dev, dtype=bf16), wp_r_t=torch.empty(c, c, device=dev, dtype=bf16),
                 w2n_t=torch.empty(c, c, device=dev, dtype=bf16), wd_t8=torch.zeros(c, 8, device=dev, dtype=bf16))
        block.__dict__["_bwdpk_pose"] = p
    # refreshed in place after every optimiser step (fixed addresses, see attention._Packed)
    p["wp_x_t"].copy_(wp[:, :c].t())
    p["wp_r_t"].copy_(wp[:, c:].t())
    p["w2n_t"].copy_(pk["w2"].t())
    p["wd_t8"][:, : pk["wd"].shape[0]].copy_(pk["wd"].t())
    p["stale"] = False
    return p


def _grad_buf(param: torch.nn.Parameter) -> torch.Tensor:
    """fp32 gradient storage of a trainable parameter (a view of the flat gradient buffer when the
    optimiser has flattened the parameters; allocated on first use otherwise)."""
    if param.grad is None:
        param.grad = torch.zeros_like(param, dtype=f32)
    return param.grad


WGRAD_TN = os.environ.get("CD360_WGRAD_TN", "1") != "0"   # 0: transposed copies + the K-major GEMM (A/B runs)


def _wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None):
    """out[N, K] (fp32, may be a strided view) = dy^T x for dy bf16/fp32 [M, N], x bf16 [M, K]: the
    contraction over the M token rows reads both activations in place (MN-major tcgen05 operands)."""
    ok = lambda t: t.stride(1) == 1 and (t.stride(0) & 7) == 0 and (t.shape[1] & 7) == 0 and (t.data_ptr() & 15) == 0
    if WGRAD_TN and x.dtype == torch.bfloat16 and ok(x):
        if dy.dtype != torch.bfloat16:
            dy = ops.cast_bf16(dy.contiguous())
        if ok(dy):
            return ops.gemm_tn(dy, x, out=out)
    return ops.gemm(ops.transpose_to_bf16(dy), ops.transpose_to_bf16(x), out=out, out_fp32=True)


# ------------------------------------------------------------------------------------------------
# FeatureNeRF: reference_attn forward (taped) and backward
# ------------------------------------------------------------------------------------------------
def nerf_bins(nerf, hw: int, dev, jitter: Optional[dict]):
    """(xy [hw,2], depths [hw,d], dists [hw,d]).  With `jitter` = {"xy_rand": (rx [res+1], ry [res+1]),
    "t_rand": [hw, d+1]} (one such dict per pose block and call) the stratified training-time sampling of the reference is reproduced from
    the INJECTED uniform variates (get_patch_raybundle, utils_cameraray.py:111-140;
    Raymarcher.stratified_sampling, nerfsd_pytorch3d.py:317-325) — parity needs the same random
    numbers, not the same generator."""
    res = int(math.sqrt(hw))
    rm = nerf.raymarcher
    if not jitter:
        depths, dists = rm.bins(hw, dev)
        return patch_ray_xy(res, dev), depths, dists
    if "bins" in jitter:      # already on the device (graph-replayed step: static buffers)
        return jitter["bins"]

    def jittered_positions(r):
        edges = torch.linspace(1, -1, res + 1, dtype=f32)
        center = (edges[1:] + edges[:-1]) / 2.0
        upper = torch.cat([center, edges[-1:]], -1)
        lower = torch.cat([edges[:1], center], -1)
        return (lower + (upper - lower) * r.float().cpu())[:-1]

    xr, yr = jitter["xy_rand"]
    hpos, vpos = jittered_positions(xr), jittered_positions(yr)
    xs = hpos[None, :].expand(res, res)
    ys = vpos[:, None].expand(res, res)
    xy = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).contiguous().to(dev)
    lower = rm.lengths_lower.to(device=dev, dtype=f32)
    upper = rm.lengths_upper.to(device=dev, dtype=f32)
    jit = lower[None] + (upper - lower)[None] * jitter["t_rand"].to(device=dev, dtype=f32)
    depths = ((jit[:, :-1] + jit[:, 1:]) / 2.0).contiguous()
    dists = (jit[:, 1:] - jit[:, :-1]).contiguous()
    return xy, depths, dists


def nerf_forward(block: BasicTransformerBlock, cams, xref_tok, n, kv, nkv, batch, hw, jitter=None):
    """reference_attn (attention.py:571-598) keeping what the backward needs.
    Returns (rendered bf16 [batch*hw, c], (fg, alphas, rgb), saved)."""
    nerf = block.pose_featurenerf
    if not nerf.rgb_predict:
        raise NotImplementedError("training path is built for rgb_predict=True (the shipped config)")
    pk = nerf.model.packed()
    c, d = pk["c"], nerf.raymarcher.num_samples
    res = int(math.sqrt(hw))
    dev = xref_tok.device
    xy, depths, dists = nerf_bins(nerf, hw, dev, jitter)
    # reference padding masks (nerfsd_pytorch3d.py:61-70); the masked tokens are also what dWg reads
    xref_tok = nerf.apply_mask_ref(xref_tok, block.__dict__.get("_mask_ref"), batch, n, hw)
    g = ops.gemm(xref_tok, pk["wg"])
    pe, gidx, gwgt, vlogit = ops.nerf_points(cams, xy, depths, pk["wnv_geo"], pk["bnv"], batch, n, res, d, KPE)
    hpre = ops.gemm(pe, pk["w1p"], bias=pk["b1"])
    s, _ = ops.nerf_combine(g, hpre, gidx, gwgt, vlogit, batch, n, hw, d, c)
    final = ops.gemm(s, pk["w2"], bias=pk["b2"])
    raw = ops.gemm(final, pk["wd"], out_fp32=True)                      # [P', 4] (rgb 3, sigma 1)
    a2 = block.attn2
    p2 = a2.packed()
    inner = a2.heads * a2.dim_head
    q = ops.gemm(block.norm2.tokens(final), p2["wq"])
    k, v = kv[:, :inner], kv[:, inner:2 * inner]
    att = ops.attention(q, k, v, batch, a2.heads, hw * d, nkv, ldq=inner, ldk=kv.stride(0), ldv=kv.stride(0))
    feats2 = ops.gemm(att, p2["wo"], bias=p2["bo"], residual=final)
    rendered, fg, alphas, rgb = ops.nerf_volrender(feats2, raw, dists, batch, hw, d, c)
    saved = NS(cams=cams, xref=xref_tok, n=n, g=g, pe=pe, gidx=gidx, gwgt=gwgt, vlogit=vlogit, hpre=hpre,
               s=s, final=final, raw=raw, q=q, k=k, v=v, att=att, feats2=feats2, dists=dists, nkv=nkv,
               batch=batch, hw=hw, d=d, c=c)
    return rendered, (fg, alphas, rgb), saved


def nerf_backward(block: BasicTransformerBlock, sv, d_rendered, daux):
    """Gradients of the FeatureNeRF weights of `block` from d(rendered) and the gradients of the
    supervised outputs daux = (dfg [b,hw], dalphas [b,hw,d], drgb [b,hw,3]) (any may be None)."""
    bp = bwd_pack(block)
    pp = pose_bwd_pack(block)
    model = block.pose_featurenerf.model
    a2 = block.attn2
    b, hw, d, c, n = sv.batch, sv.hw, sv.d, sv.c, sv.n
    dfg, dal, drgb = daux if daux is not None else (None, None, None)
    dfeats2, draw8 = ops.nerf_volrender_bwd(sv.feats2, sv.raw, sv.dists, d_rendered, dfg, dal, drgb, b, hw, d, c)
    # feats2 = final + to_out(attn(to_q(LN2 final), K, V))   (the block's own norm2 / attn2, frozen)
    da = ops.gemm(dfeats2, bp["wo2_t"])
    dq = torch.empty_like(sv.q)
    dkv = _BWD.get("dkv_nerf")          # conditioning gradients: this attn2 call also reads the text K / V
    if dkv is None:
        ops.attention_bwd(sv.q, sv.k, sv.v, sv.att, da, b, a2.heads, hw * d, sv.nkv, dq=dq)
    else:
        off, _ = block.__dict__["_kv_slice"]
        inner = a2.heads * a2.dim_head
        ops.attention_bwd(sv.q, sv.k, sv.v, sv.att, da, b, a2.heads, hw * d, sv.nkv, dq=dq,
                          dk=dkv[:, off:off + inner], dv=dkv[:, off + inner:off + 2 * inner])
    dfn = ops.gemm(dq, bp["wq2_t"])
    dfinal = ops.layernorm_bwd(sv.final, bp["g2"], dfn, add=dfeats2, eps=block.norm2.eps)
    # raw = final Wd^T
    dfinal = ops.gemm(draw8, pp["wd_t8"], residual=dfinal, out=dfinal)
    dwd = _wgrad(draw8, sv.final)                                                  # [8, c]
    _grad_buf(model.decoder.weight).copy_(dwd[: model.decoder.weight.shape[0]])
    # final = S W2^T + b2
    _wgrad(dfinal, sv.s, _grad_buf(model.plane_coefs[2].weight))
    ops.colsum(dfinal, out=_grad_buf(model.plane_coefs[2].bias).zero_())
    ds = ops.gemm(dfinal, pp["w2n_t"])
    # gather / SiLU / view softmax
    dhpre, dlogit, dg = ops.nerf_combine_bwd(sv.g, sv.hpre, sv.gidx, sv.gwgt, sv.vlogit, ds, b, n, hw, d, c)
    # hpre = pe W1p^T + b1  |  G = xref [W1f ; w_nv_f]^T
    dw1 = _grad_buf(model.plane_coefs[0].weight)                                   # [c, c + 198]
    dw1p = _wgrad(dhpre, sv.pe)                                                    # [c, KPE]
    dw1[:, c:].copy_(dw1p[:, :198])
    ops.colsum(dhpre, out=_grad_buf(model.plane_coefs[0].bias).zero_())
    dwg = _wgrad(dg[:, : c + 8], sv.xref)                                          # [c+8, c]
    dw1[:, :c].copy_(dwg[:c])
    dnv = _grad_buf(model.nviews.weight)                                           # [1, c + 198]
    dnv[0, :c].copy_(dwg[c])
    dnv[0, c:].copy_(ops.nerf_nviews_geo_bwd(sv.cams, dlogit, b, n))
    _grad_buf(model.nviews.bias).zero_()                                           # sum_v dlogit_v == 0


# ------------------------------------------------------------------------------------------------
# BasicTransformerBlock
# ------------------------------------------------------------------------------------------------
def block_forward(block: BasicTransformerBlock, x, batch, n, kv, nctx, cams=None, jitter=None, stats=None):
    """Taped `BasicTransformerBlock._forward` (attention.py:600-637) in token layout.  Residual
    updates are out of place: every saved tensor stays valid until the backward.
    stats: fp32 [M, c/64, 2] row moments of x written by the GEMM that produced x -> the three
    LayerNorms are folded into the consuming GEMMs' epilogues (attention.BasicTransformerBlock.ln_packed,
    same algebra as the sampling path); None -> stand-alone LayerNorm kernels.  Returns
    (x4, aux, saved, stats of x4 | None)."""
    a1, a2 = block.attn1, block.attn2
    p1, p2 = a1.packed(), a2.packed()
    inner = a1.heads * a1.dim_head
    M, c = x.shape
    fused = stats is not None
    eps = block.norm1.eps
    new_stats = (lambda: torch.empty((M, c // 64, 2), device=x.device, dtype=f32)) if fused else (lambda: None)
    lp = block.ln_packed() if fused else None
    if fused:
        qkv = ops.gemm(x, lp["wqkv"], bias=lp["bqkv"], ln_stats=stats, ln_colsum=lp["cqkv"], ln_eps=eps)
    else:
        qkv = ops.gemm(block.norm1.tokens(x), p1["wqkv"])
    att1 = ops.attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], batch, a1.heads, n, n,
                         ldq=3 * inner, ldk=3 * inner, ldv=3 * inner)
    st1 = new_stats()
    x1 = ops.gemm(att1, p1["wo"], bias=p1["bo"], residual=x, stats_out=st1)
    if fused:
        q2 = ops.gemm(x1, lp["wq2"], bias=lp["bq2"], ln_stats=st1, ln_colsum=lp["cq2"], ln_eps=eps)
    else:
        q2 = ops.gemm(block.norm2.tokens(x1), p2["wq"])
    k2, v2 = kv[:, :inner], kv[:, inner:2 * inner]
    att2 = ops.attention(q2, k2, v2, batch, a2.heads, n, nctx, ldq=inner, ldk=kv.stride(0), ldv=kv.stride(0))
    st2 = new_stats()
    x2 = ops.gemm(att2, p2["wo"], bias=p2["bo"], residual=x1, stats_out=st2)
    sv = NS(x0=x, qkv=qkv, att1=att1, x1=x1, q2=q2, k2=k2, v2=v2, att2=att2, x2=x2, batch=batch, n=n,
            nctx=nctx, inner=inner, nerf=None, rendered=None)
    aux = None
    x3, st3 = x2, st2
    if block.image_cross and cams is not None:
        pre = _FWD.get(id(block))
        if pre is not None:      # FeatureNeRF already enqueued on its own stream (unet_forward)
            rendered, aux, sv.nerf, ev = pre
            torch.cuda.current_stream(x.device).wait_event(ev)
        else:
            xref_tok = block.context_ref_tokens(batch)
            n_views = block._ctxref_cache[2]
            rendered, aux, sv.nerf = nerf_forward(block, cams, xref_tok, n_views, kv, nctx, batch, n,
                                                  next(jitter) if jitter is not None else None)
        sv.rendered = rendered
        st3 = new_stats()
        x3 = block.pose_emb_layers.tokens(x2, a1=rendered, stats_out=st3)
    sv.x3 = x3
    # feed-forward.  The GEGLU pre-activation is KEPT (bf16 [M, 8c], packed column order) instead of being
    # recomputed in the backward: one elementwise launch here replaces a LayerNorm + an [M, 8c, c] GEMM there.
    ffp = block.ff.net[0].packed()
    if fused:
        raw = ops.gemm(x3, lp["wff"], bias=lp["bff"], ln_stats=st3, ln_colsum=lp["cff"], ln_eps=eps)
    else:
        raw = ops.gemm(block.norm3.tokens(x3), ffp["w"], bias=ffp["b"])
    sv.ff_raw = raw
    h = ops.geglu_fwd(raw, ops.geglu_pack_block(raw.shape[1]))
    st4 = new_stats()
    x4 = block.ff.net[2].tokens(h, residual=x3, stats_out=st4)
    return x4, aux, sv, st4


def block_backward(block: BasicTransformerBlock, sv, g, daux, stop_here: bool):
    """g = dL/d(block output) bf16 [M, c] -> dL/d(block input), or None when the walk stops at this
    block's pose layers (`stop_here`: nothing upstream is trainable)."""
    bp = bwd_pack(block)
    ff = block.ff
    blk = ops.geglu_pack_block(ff.net[0].proj.out_features)
    # ---- feed-forward: x4 = x3 + W2 geglu(Wff LN3(x3) + bff) + b2
    dh = ops.gemm(g, bp["w2_t"])
    draw = ops.geglu_bwd(sv.ff_raw, dh, blk)          # pre-activation kept by the forward (packed column order)
    del dh
    g = ops.layernorm_bwd(sv.x3, bp["g3"], ops.gemm(draw, bp["wff_t"]), add=g, eps=block.norm3.eps)
    del draw
    # ---- pose_emb_layers: x3 = [x2 | rendered] Wp^T
    if sv.nerf is not None:
        c = sv.x2.shape[1]
        pp = pose_bwd_pack(block)          # (refreshed on the current stream, before the fork)
        # Everything that only produces WEIGHT gradients of this block — the pose_emb_layers dW GEMMs
        # and the whole FeatureNeRF backward (its inputs are the no-grad reference tokens, so nothing
        # flows on from it) — is off the critical dX chain: it runs on the side stream, concurrently
        # with the rest of the walk, and is joined at the end of unet_backward.
        side = _BWD.get("side")
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(g.device))
            _BWD["keep"].append((g, sv, daux))   # read on the side stream: alive until the join
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            dwp = _grad_buf(block.pose_emb_layers.weight)                       # [c, 2c]
            _wgrad(g, sv.x2, dwp[:, :c])
            _wgrad(g, sv.rendered, dwp[:, c:])
            nerf_backward(block, sv.nerf, ops.gemm(g, pp["wp_r_t"]), daux)
            ready = block.__dict__.get("_grads_ready")     # data-parallel: start this block's all-reduce now
            if ready is not None:
                ready()
        if stop_here:
            return None
        g = ops.gemm(g, pp["wp_x_t"])
    # ---- text cross-attention (K / V: projections of the context; their gradient only when the
    #      conditioning gradients are asked for)
    a1, a2 = block.attn1, block.attn2
    da = ops.gemm(g, bp["wo2_t"])
    dq = torch.empty_like(sv.q2)
    dkv = _BWD.get("dkv")
    if dkv is None:
        ops.attention_bwd(sv.q2, sv.k2, sv.v2, sv.att2, da, sv.batch, a2.heads, sv.n, sv.nctx, dq=dq)
    else:
        off, _ = block.__dict__["_kv_slice"]
        ops.attention_bwd(sv.q2, sv.k2, sv.v2, sv.att2, da, sv.batch, a2.heads, sv.n, sv.nctx, dq=dq,
                          dk=dkv[:, off:off + sv.inner], dv=dkv[:, off + sv.inner:off + 2 * sv.inner])
    g = ops.layernorm_bwd(sv.x1, bp["g2"], ops.gemm(dq, bp["wq2_t"]), add=g, eps=block.norm2.eps)
    # ---- self-attention
    inner = sv.inner
    da = ops.gemm(g, bp["wo1_t"])
    dqkv = torch.empty_like(sv.qkv)
    ops.attention_bwd(sv.qkv[:, :inner], sv.qkv[:, inner:2 * inner], sv.qkv[:, 2 * inner:], sv.att1, da,
                      sv.batch, a1.heads, sv.n, sv.n, dq=dqkv[:, :inner], dk=dqkv[:, inner:2 * inner],
                      dv=dqkv[:, 2 * inner:])
    return ops.layernorm_bwd(sv.x0, bp["g1"], ops.gemm(dqkv, bp["wqkv_t"]), add=g, eps=block.norm1.eps)


# ------------------------------------------------------------------------------------------------
# SpatialTransformer / ResBlock / resampling
# ------------------------------------------------------------------------------------------------
def st_forward(st: SpatialTransformer, x, batch, hw, nctx, kv_all, cams, aux_out, jitter=None):
    p = st.packed()
    xn = ops.groupnorm(x, p["g"], p["b"], batch, hw, eps=st.norm.eps, silu=False)
    from .attention import LN_FUSED
    stats = None
    if LN_FUSED and st.in_channels % 64 == 0:   # proj_in's epilogue emits the row moments of the first LayerNorm
        stats = torch.empty((xn.shape[0], st.in_channels // 64, 2), device=xn.device, dtype=f32)
        h = st.proj_in.tokens(xn, stats_out=stats)
    else:
        h = st.proj_in.tokens(xn)
    saved = []
    for i, block in enumerate(st.transformer_blocks):
        use_pose = st.image_cross and (i % st.poscontrol_interval == 0)
        off, width = block.__dict__["_kv_slice"]
        h, aux, sv, stats = block_forward(block, h, batch, hw, kv_all[:, off:off + width], nctx,
                                          cams if use_pose else None, jitter, stats=stats)
        if aux is not None:
            aux_out.append((block, aux))
        saved.append(sv)
    out = st.proj_out.tokens(h, residual=x)
    return out, NS(x=x, blocks=saved, batch=batch, hw=hw)


def st_backward(st: SpatialTransformer, sv, g, daux_of, first_pose_block):
    bp = bwd_pack(st)
    p = st.packed()
    dh = ops.gemm(g, bp["wout_t"])
    for block, bsv in zip(reversed(list(st.transformer_blocks)), reversed(sv.blocks)):
        dh = block_backward(block, bsv, dh, daux_of.get(id(block)), stop_here=block is first_pose_block)
        if dh is None:
            return None
    dxn = ops.gemm(dh, bp["win_t"])
    dx, _ = ops.groupnorm_bwd(sv.x, p["g"], p["b"], dxn, sv.batch, sv.hw, add0=g, eps=st.norm.eps, silu=False)
    return dx


def res_forward(rb, x, batch, h, w, emb_out, skip=None, emb_slice=None):
    p = rb.packed()
    hw = h * w
    hn = ops.groupnorm(x, p["g1"], p["b1"], batch, hw, x1=skip, eps=rb.in_layers[0].eps, silu=True)
    h1 = ops.conv3x3(hn, p["w1"], batch, h, w, bias=p["cb1"], row_bias=emb_out)
    hn2 = ops.groupnorm(h1, p["g2"], p["b2"], batch, hw, eps=rb.out_layers[0].eps, silu=True)
    xs = ops.gemm(x, p["ws"], bias=p["bs"], a1=skip) if "ws" in p else x
    out = ops.conv3x3(hn2, p["w2"], batch, h, w, bias=p["cb2"], residual=xs)
    return out, NS(x=x, skip=skip, h1=h1, batch=batch, h=h, w=w, emb_slice=emb_slice)


def res_backward(rb, sv, g):
    """-> (dL/dx, dL/dskip | None)"""
    bp = bwd_pack(rb)
    p = rb.packed()
    b, h, w = sv.batch, sv.h, sv.w
    hw = h * w
    dhn2 = ops.conv3x3(g, bp["w2"], b, h, w)
    dh1, _ = ops.groupnorm_bwd(sv.h1, p["g2"], p["b2"], dhn2, b, hw, eps=rb.out_layers[0].eps, silu=True)
    demb = _BWD.get("demb")
    if demb is not None:     # h1 = conv(...) + emb_out[:, :, None, None]: per-image sum over the pixels
        off, n = sv.emb_slice
        for i in range(b):
            ops.colsum(dh1[i * hw:(i + 1) * hw], out=demb[i, off:off + n])
    dhn = ops.conv3x3(dh1, bp["w1"], b, h, w)
    c0 = sv.x.shape[1]
    if "ws_t" in bp:
        dxs = ops.gemm(g, bp["ws_t"])
        add0, add1 = dxs[:, :c0], (dxs[:, c0:] if sv.skip is not None else None)
    else:
        add0, add1 = g, None
    return ops.groupnorm_bwd(sv.x, p["g1"], p["b1"], dhn, b, hw, x1=sv.skip, add0=add0, add1=add1,
                             eps=rb.in_layers[0].eps, silu=True)


def down_backward(layer, g, batch, h, w):
    """Downsample (stride-2 conv through im2col + GEMM); h, w = INPUT size."""
    bp = bwd_pack(layer)
    dcol = ops.gemm(g, bp["w_t"])
    return ops.col2im3x3_s2(dcol, batch, h, w, layer.channels)


def up_backward(layer, g, batch, h, w):
    """Upsample = nearest x2 + conv3x3; h, w = INPUT size."""
    bp = bwd_pack(layer)
    dup = ops.conv3x3(g, bp["w"], batch, 2 * h, 2 * w)
    return ops.upsample_nearest2x_bwd(dup, batch, h, w)


# ------------------------------------------------------------------------------------------------
# UNet
# ------------------------------------------------------------------------------------------------
def unet_forward(unet, x, timesteps, context, y, pose, in_scale=None, jitter=None, cond_grad=False):
    """Taped main-stream forward (pose blocks read the live reference-stream tokens installed by the
    caller).  Returns (eps fp32 tokens [B*L*L, 4], aux list [(block, (fg, alphas, rgb))], tape).
    Tape entries, in forward order: ("layer", kind, module, saved), ("push", i) = h became skip
    tensor hs[i] (openaimodel.py:1055-1071), ("pop", i) = the next ResBlock consumed hs[i] (:1074)."""
    from .diffusionmodules import openaimodel as U
    from .attention import to_tokens
    from .utils_cameraray import pack_pose
    from ..._lib import ACT_SILU

    p = unet.packed()
    b, cin, hh, ww = x.shape
    dev = x.device
    t_emb = ops.timestep_embedding(timesteps.to(device=dev, dtype=f32).contiguous(), unet.model_channels)
    e1 = ops.small_linear(t_emb, p["te0w"], p["te0b"], act_out=ACT_SILU)
    emb = ops.small_linear(e1, p["te2w"], p["te2b"])
    # (label_emb's SiLU is applied on the NEXT layer's input so that its pre-activation is kept for the
    # backward to dL/d(vector); same arithmetic as act_out=SILU here)
    l1p = ops.small_linear(y.float().contiguous(), p["le0w"], p["le0b"])
    emb = ops.small_linear(l1p, p["le2w"], p["le2b"], add=emb, act_in=ACT_SILU)
    emb_all = ops.small_linear(emb, p["embw"], p["embb"], act_in=ACT_SILU)
    ctx_tok, nctx = to_tokens(context), context.shape[1]
    cams = pack_pose(pose, dev) if pose is not None else None
    kv_all = ops.gemm(ctx_tok, p["kvw"])
    aux: list = []
    tape: list = []
    jitter = iter(jitter) if jitter else None   # per-pose-block variates, consumed in execution order

    def run(layers, h, hh, ww, skip=None):
        for layer in layers:
            if isinstance(layer, U.ResBlock):
                off, n = p["emb_off"][id(layer)]
                h, sv = res_forward(layer, h, b, hh, ww, emb_all[:, off:off + n], skip=skip, emb_slice=(off, n))
                skip = None
                tape.append(("layer", "res", layer, sv))
            elif isinstance(layer, SpatialTransformer):
                h, sv = st_forward(layer, h, b, hh * ww, nctx, kv_all, cams, aux, jitter)
                tape.append(("layer", "st", layer, sv))
            elif isinstance(layer, U.Downsample):
                tape.append(("layer", "down", layer, NS(h=hh, w=ww)))
                h = layer.tokens(h, b, hh, ww)
                hh, ww = hh // 2, ww // 2
            elif isinstance(layer, U.Upsample):
                tape.append(("layer", "up", layer, NS(h=hh, w=ww)))
                h = layer.tokens(h, b, hh, ww)
                hh, ww = hh * 2, ww * 2
            else:
                raise TypeError(type(layer))
        return h, hh, ww

    # FeatureNeRF of every pose block depends on the reference tokens, the cameras and the weights —
    # not on the main stream's activations: enqueue all of them on their own stream (each waits for
    # the event of its block's reference tokens) so they run beside the main stream, which only waits
    # for a block's event when it reaches that block's pose_emb_layers.
    from .diffusionmodules.openaimodel import OVERLAP_REF_STREAM
    nerf_stream = None
    _FWD.clear()
    if cams is not None and x.is_cuda and OVERLAP_REF_STREAM and unet.__dict__.get("_packs_warm"):
        nerf_stream = unet.__dict__.get("_nerf_stream")
        if nerf_stream is None or nerf_stream.device != dev:
            nerf_stream = torch.cuda.Stream(device=dev)
            unet.__dict__["_nerf_stream"] = nerf_stream
        nerf_stream.wait_stream(torch.cuda.current_stream(dev))      # kv_all, cams
        with torch.cuda.stream(nerf_stream):
            for block in pose_blocks_in_order(unet):
                c = block.pose_emb_layers.weight.shape[0]
                res = x.shape[-1] // (c // unet.model_channels)
                off, width = block.__dict__["_kv_slice"]
                xref_tok = block.context_ref_tokens(b)                # waits for the reference stream's event
                n_views = block._ctxref_cache[2]
                rendered, a, saved = nerf_forward(block, cams, xref_tok, n_views, kv_all[:, off:off + width], nctx,
                                                  b, res * res, next(jitter) if jitter is not None else None)
                ev = torch.cuda.Event()
                ev.record(nerf_stream)
                _FWD[id(block)] = (rendered, a, saved, ev)
    try:
        col = ops.im2col3x3_nchw(x.float().contiguous(), 64, scale=in_scale, batch=b)
        h = ops.gemm(col, p["cin_w"], bias=p["cin_b"])
        hs = [h]
        tape.append(("push", 0))
        for block in list(unet.input_blocks)[1:]:
            h, hh, ww = run(block, h, hh, ww)
            tape.append(("push", len(hs)))
            hs.append(h)
        h, hh, ww = run(unet.middle_block, h, hh, ww)
        for block in unet.output_blocks:
            tape.append(("pop", len(hs) - 1))
            h, hh, ww = run(block, h, hh, ww, skip=hs.pop())
        hn = ops.groupnorm(h, p["og"], p["ob"], b, hh * ww, eps=unet.out[0].eps, silu=True)
        eps = ops.conv3x3(hn, p["cout_w"], b, hh, ww, bias=p["cout_b"], out_fp32=True)
    finally:
        if nerf_stream is not None:
            torch.cuda.current_stream(dev).wait_stream(nerf_stream)
        _FWD.clear()
    cond = NS(l1p=l1p, emb=emb, nctx=nctx, kv_width=kv_all.shape[1], emb_width=emb_all.shape[1],
              ctx_dim=ctx_tok.shape[1]) if cond_grad else None
    return eps, aux, NS(tape=tape, h_last=h, batch=b, hh=hh, ww=ww, cond=cond)


def pose_blocks_in_order(unet) -> List[BasicTransformerBlock]:
    """Pose blocks in the order the forward pass executes them."""
    return [m for blk in list(unet.input_blocks) + [unet.middle_block] + list(unet.output_blocks)
            for layer in blk if isinstance(layer, SpatialTransformer) and layer.image_cross
            for i, m in enumerate(layer.transformer_blocks) if m.image_cross and i % layer.poscontrol_interval == 0]


def first_pose_block(unet) -> Optional[BasicTransformerBlock]:
    """The pose block that runs first in the forward pass: the backward walk ends there."""
    order = pose_blocks_in_order(unet)
    return order[0] if order else None


def unet_backward(unet, fw, deps, daux_of: Dict[int, tuple]):
    """deps: bf16 [B*L*L, 64] gradient of the loss w.r.t. the UNet output tokens (columns >= 4
    zero); daux_of: {id(pose block): (dfg, dalphas, drgb)}.  Writes the gradients of every pose
    parameter into its `.grad` (fp32).  When the forward was taped with cond_grad=True, returns
    {"crossattn": fp32 [b, nctx, ctx_dim], "vector": fp32 [b, adm]} (else None)."""
    from .diffusionmodules.openaimodel import OVERLAP_REF_STREAM
    side = None
    cg = fw.cond
    dev = deps.device
    if cg is not None:
        unet.__dict__["_cond_grad"] = True
        b = fw.batch
        # every transformer block writes its own K|V column slice (all 70 are visited); the FeatureNeRF
        # branch runs on the side stream and accumulates into a buffer of its own, added after the join
        _BWD["dkv"] = torch.empty(b * cg.nctx, cg.kv_width, device=dev, dtype=bf16)
        _BWD["dkv_nerf"] = torch.zeros(b * cg.nctx, cg.kv_width, device=dev, dtype=bf16)
        _BWD["demb"] = torch.zeros(b, cg.emb_width, device=dev, dtype=f32)
    else:
        _BWD["dkv"] = _BWD["dkv_nerf"] = _BWD["demb"] = None
    if deps.is_cuda and OVERLAP_REF_STREAM and unet.__dict__.get("_packs_warm") and unet.__dict__.get("_bwd_warm"):
        side = unet.__dict__.get("_side_stream")
        if side is None or side.device != deps.device:
            side = torch.cuda.Stream(device=deps.device)
            unet.__dict__["_side_stream"] = side
    _BWD["side"], _BWD["keep"] = side, []
    try:
        _unet_backward(unet, fw, deps, daux_of)
        unet.__dict__["_bwd_warm"] = True     # the lazily built backward packs exist from now on
    finally:
        if side is not None:
            torch.cuda.current_stream(deps.device).wait_stream(side)
        _BWD["side"], _BWD["keep"] = None, []
        dkv, dkv_nerf, demb = _BWD.get("dkv"), _BWD.get("dkv_nerf"), _BWD.get("demb")
        _BWD["dkv"] = _BWD["dkv_nerf"] = _BWD["demb"] = None
    if cg is None:
        return None
    return _cond_backward(unet, cg, fw.batch, dkv, dkv_nerf, demb)


def _cond_backward(unet, cg, b, dkv, dkv_nerf, demb):
    """dL/d(context) and dL/d(vector) from the collected dK|dV and emb_layers gradients."""
    bp = bwd_pack(unet)
    if "kvw_t" not in bp:      # the pack was built before the first conditioning-gradient request
        unet.__dict__.pop("_bwdpk", None)
        bp = bwd_pack(unet)
    ops.add_bf16(dkv, dkv_nerf, out=dkv)
    dctx = ops.gemm(dkv, bp["kvw_t"], out_fp32=True)                     # K|V = ctx [Wk ; Wv]^T  (to_k / to_v, no bias)
    # emb_out = Linear(SiLU(emb));  emb = time_embed(t) + Linear(SiLU(Linear(y)))   (openaimodel.py:679-713, 1026-1031)
    # (K = sum of the 17 ResBlocks' widths exceeds the small-linear kernel's smem-resident row: tensor-core GEMM, split-K)
    d_act = ops.gemm(ops.cast_bf16(demb), bp["embw_t"], out_fp32=True)
    d_emb = ops.silu_bwd(cg.emb, d_act)
    d_l1 = ops.small_linear(d_emb, bp["le2w_t"])
    d_l1p = ops.silu_bwd(cg.l1p, d_l1)
    dy = ops.small_linear(d_l1p, bp["le0w_t"])
    return {"crossattn": dctx.view(b, cg.nctx, cg.ctx_dim), "vector": dy}


def _unet_backward(unet, fw, deps, daux_of):
    p = unet.packed()
    bp = bwd_pack(unet)
    b = fw.batch
    stop = first_pose_block(unet) if fw.cond is None else None    # conditioning gradients need the whole walk
    g = ops.conv3x3(deps, bp["cout"], b, fw.hh, fw.ww)
    g, _ = ops.groupnorm_bwd(fw.h_last, p["og"], p["ob"], g, b, fw.hh * fw.ww, eps=unet.out[0].eps, silu=True)
    skip_grads: Dict[int, torch.Tensor] = {}
    pending_skip = None
    trace = unet.__dict__.get("_grad_trace")   # tests: list collecting (tag, activation gradient)
    n_out = len(unet.output_blocks)
    if trace is not None:
        trace.append((f"out{n_out - 1}", g.clone()))
    for entry in reversed(fw.tape):
        if entry[0] == "push":
            sg = skip_grads.pop(entry[1], None)
            if sg is not None:
                g = ops.add_bf16(g, sg)
            continue
        if entry[0] == "pop":
            skip_grads[entry[1]] = pending_skip
            pending_skip = None
            if trace is not None:   # g = gradient of the previous block's output
                n_out -= 1
                trace.append((f"out{n_out - 1}" if n_out > 0 else "middle", g.clone()))
            continue
        _, kind, layer, sv = entry
        if kind == "res":
            g, dskip = res_backward(layer, sv, g)
            if dskip is not None:
                pending_skip = dskip
        elif kind == "st":
            g = st_backward(layer, sv, g, daux_of, stop)
            if g is None:
                return
        elif kind == "down":
            g = down_backward(layer, g, b, sv.h, sv.w)
        elif kind == "up":
            g = up_backward(layer, g, b, sv.h, sv.w)
