"""Transformer / attention surface of the reference (sgm/modules/attention.py) on the sm_100a
kernels: GEGLU, FeedForward, MemoryEfficientCrossAttention, BasicTransformerBlock (with the
FeatureNeRF injection) and SpatialTransformer.

Drop-in contract: class names, constructor kwargs, attribute names
(`attn1/attn2/ff/norm1-3/pose_emb_layers/pose_featurenerf/renderer/rendered_feat/references`,
`norm/proj_in/transformer_blocks/proj_out/use_linear/image_cross/poscontrol_interval`) and
state-dict keys equal the reference's.  Two call paths exist per module:

  * `forward(...)` — the reference's tensor contract ([B, N, c] / [B, c, H, W] fp32 in and out),
    for callers that address sub-modules directly (sample.py's patched forwards do);
  * `tokens(...)`   — bf16 token layout [B*N, c] straight between kernels, what UNetModel uses.

Inference semantics built in (what sample.py:33-136 monkey-patches onto the reference): pose
blocks read their reference tokens from the `references` buffer + `choices`, and cache
`rendered_feat` after the first step until `clear_rendered_feat()`.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn

from ... import ops
from ..._lib import ACT_NONE
from ..prepack import pack_geglu
from .nerfsd_pytorch3d import NerfSDModule, VolRender
from .utils_cameraray import pack_pose

bf16 = torch.bfloat16
# LayerNorms folded into the neighbouring GEMM epilogues on the token fast path (removes 210
# launches per step).  Measured on B200 (profiles/README_r01.md): a loss while the epilogue fetched
# bias / column sums from global memory (29.66 vs 29.37 ms / step), a gain since they are staged in
# shared memory (26.1 vs 26.5 ms), so it is ON by default; CD360_LN_FUSED=0 disables it
# (parity-tested either way).
import os as _os
LN_FUSED = _os.environ.get("CD360_LN_FUSED", "1") != "0"


def to_tokens(x: torch.Tensor) -> torch.Tensor:
    """[..., c] float tensor -> contiguous bf16 [rows, c]."""
    t = x.reshape(-1, x.shape[-1])
    if t.dtype == bf16:
        return t.contiguous()
    return ops.cast_bf16(t.float().contiguous())


class _Packed:
    """Lazily built bf16 / re-laid-out copies of a module's parameters (kernel operand layout).

    Two ways to invalidate: `invalidate_packed()` drops the copies (rebuilt into NEW tensors on next
    use); `mark_stale()` keeps the tensors and has the next `packed()` call refresh them IN PLACE
    (`_refresh`) — what the training loop uses after every optimiser step, so that the addresses a
    captured CUDA graph (GraphedTrainStep) and the eager step read never change and the refresh
    itself is part of the graph."""

    def packed(self):
        dev = next(self.parameters()).device
        p = self.__dict__.get("_pk")
        if p is None or p["dev"] != dev:
            p = self._pack(dev)
            p["dev"] = dev
            self.__dict__["_pk"] = p
            self.__dict__["_pk_stale"] = False
        elif self.__dict__.get("_pk_stale"):
            self._refresh(p)
            self.__dict__["_pk_stale"] = False
        return p

    def _refresh(self, p):
        fresh = self._pack(p["dev"])
        for k, v in fresh.items():
            if torch.is_tensor(v) and v.data_ptr() != p[k].data_ptr():
                p[k].copy_(v)

    def invalidate_packed(self):
        self.__dict__["_pk"] = None

    def mark_stale(self):
        if self.__dict__.get("_pk") is not None:
            self.__dict__["_pk_stale"] = True


def invalidate_all_packed(root: nn.Module, only_trainable: bool = False, keep_buffers: bool = True):
    """Make the packed operand copies follow the parameters again.  only_trainable: just those of the
    pose weights (after an optimiser step; the frozen packs stay) — by default refreshed in place
    (`keep_buffers`), see _Packed; otherwise every pack is dropped and rebuilt."""
    if not only_trainable:
        root.__dict__["_packs_warm"] = False   # UNetModel.forward_train: next call builds packs serialised
        root.__dict__["_bwd_warm"] = False
    for m in root.modules():
        if only_trainable:
            if hasattr(m, "pose_emb_layers"):
                if keep_buffers:
                    m.pose_emb_layers.mark_stale()
                    m.pose_featurenerf.model.mark_stale()
                    if m.__dict__.get("_bwdpk_pose") is not None:
                        m.__dict__["_bwdpk_pose"]["stale"] = True
                else:
                    m.pose_emb_layers.invalidate_packed()
                    m.pose_featurenerf.model._packed = None
                    m.__dict__.pop("_bwdpk_pose", None)
            continue
        if isinstance(m, _Packed):
            m.invalidate_packed()
        if hasattr(m, "_packed"):
            m._packed = None
        m.__dict__.pop("_lnpk", None)
        m.__dict__.pop("_bwdpk", None)   # transposed / tap-flipped packs of the training backward
        m.__dict__.pop("_bwdpk_pose", None)


class Linear(nn.Linear, _Packed):
    """nn.Linear whose forward is the tcgen05 GEMM (any leading dims)."""

    def _pack(self, dev):
        return dict(w=self.weight.detach().to(bf16).contiguous(),
                    b=None if self.bias is None else self.bias.detach().float().contiguous())

    def _refresh(self, p):
        ops.cast_bf16(self.weight.detach().float().contiguous(), out=p["w"])
        if self.bias is not None and p["b"].data_ptr() != self.bias.data_ptr():
            p["b"].copy_(self.bias.detach())

    def tokens(self, x, **kw):
        p = self.packed()
        return ops.gemm(x, p["w"], bias=p["b"], **kw)

    def forward(self, x):
        y = self.tokens(to_tokens(x))
        return ops.cast_f32(y).view(*x.shape[:-1], self.out_features)


class LayerNorm(nn.LayerNorm, _Packed):
    def _pack(self, dev):
        return dict(g=self.weight.detach().float().contiguous(), b=self.bias.detach().float().contiguous())

    def tokens(self, x, out=None):
        p = self.packed()
        return ops.layernorm(x, p["g"], p["b"], eps=self.eps, out=out)

    def forward(self, x):
        return ops.cast_f32(self.tokens(to_tokens(x))).view(x.shape)


class GEGLU(nn.Module, _Packed):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def _pack(self, dev):
        w, b = pack_geglu(self.proj.weight.detach(), self.proj.bias.detach())
        return dict(w=w, b=b)

    def tokens(self, x):
        p = self.packed()
        return ops.gemm(x, p["w"], bias=p["b"], geglu=True)

    def forward(self, x):
        return ops.cast_f32(self.tokens(to_tokens(x))).view(*x.shape[:-1], self.proj.out_features // 2)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        if not glu:
            raise NotImplementedError("only the gated (GEGLU) feed-forward of the SDXL config is built")
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(dropout), Linear(inner, dim_out or dim))

    def tokens(self, xn, residual=None, out=None):
        return self.net[2].tokens(self.net[0].tokens(xn), residual=residual, out=out)

    def forward(self, x):
        return ops.cast_f32(self.tokens(to_tokens(x))).view(x.shape)


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class MemoryEfficientCrossAttention(nn.Module, _Packed):
    """softmax(QK^T/8)V with to_q/to_k/to_v/to_out projections (reference :305-425).  Heads are
    addressed in place by the attention kernel — no permute / contiguous copies."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0,
                 add_lora=False, **kwargs):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("attention kernel is specialised for head dim 64")
        if add_lora:
            raise NotImplementedError("add_lora=True is not built (shipped config: add_lora: False)")
        inner = dim_head * heads
        self.self_attention = context_dim is None
        context_dim = context_dim or query_dim
        self.heads = heads
        self.dim_head = dim_head
        self.add_lora = add_lora
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def _pack(self, dev):
        q, k, v = (m.weight.detach().to(bf16) for m in (self.to_q, self.to_k, self.to_v))
        p = dict(wo=self.to_out[0].weight.detach().to(bf16).contiguous(),
                 bo=self.to_out[0].bias.detach().float().contiguous(), wq=q.contiguous(),
                 wkv=torch.cat([k, v], 0).contiguous())
        if self.self_attention:
            p["wqkv"] = torch.cat([q, k, v], 0).contiguous()
        return p

    def project_context(self, ctx_tok):
        """K|V of the context tokens, [B*nctx, 2*inner]."""
        return ops.gemm(ctx_tok, self.packed()["wkv"])

    def tokens(self, xn, batch, n, *, kv=None, nkv=None, residual=None, out=None):
        """xn bf16 [batch*n, c] (already normalised).  kv: None -> self-attention on xn; else the
        projected context [batch*nkv, 2*inner].  Returns to_out(attn) (+ residual)."""
        p = self.packed()
        inner = self.heads * self.dim_head
        if kv is None:
            qkv = ops.gemm(xn, p["wqkv"])
            a = ops.attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], batch,
                              self.heads, n, n, ldq=3 * inner, ldk=3 * inner, ldv=3 * inner)
        else:
            q = ops.gemm(xn, p["wq"])
            # kv may be a column slice of the all-blocks context projection: row stride from the view
            a = ops.attention(q, kv[:, :inner], kv[:, inner:2 * inner], batch, self.heads, n, nkv,
                              ldq=inner, ldk=kv.stride(0), ldv=kv.stride(0))
        return ops.gemm(a, p["wo"], bias=p["bo"], residual=residual, out=out)

    def forward(self, x, context=None, mask=None, additional_tokens=None,
                n_times_crossframe_attn_in_self=0):
        if mask is not None or additional_tokens is not None or n_times_crossframe_attn_in_self:
            raise NotImplementedError("mask / additional_tokens / crossframe attention are unused on this path")
        b, n, _ = x.shape
        xt = to_tokens(x)
        if context is None and self.self_attention:
            y = self.tokens(xt, b, n)
        else:
            ctx = x if context is None else context
            kv = self.project_context(to_tokens(ctx))
            y = self.tokens(xt, b, n, kv=kv, nkv=ctx.shape[1])
        return ops.cast_f32(y).view(b, n, -1)


class BasicTransformerBlock(nn.Module):
    ATTENTION_MODES = {
        "softmax": MemoryEfficientCrossAttention,       # the reference's "softmax" class cannot be
        "softmax-xformers": MemoryEfficientCrossAttention,  # constructed; same kernel
    }

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True,
                 checkpoint=True, disable_self_attn=False, attn_mode="softmax", sdp_backend=None,
                 image_cross=False, far=2, num_samples=32, add_lora=False, rgb_predict=False,
                 mode="pixel-nerf", average=False, num_freqs=16, use_prev_weights_imp_sample=False,
                 imp_sample_next_step=False, stratified=False, imp_sampling_percent=0.9,
                 near_plane=0.0):
        super().__init__()
        assert attn_mode in self.ATTENTION_MODES
        if disable_self_attn:
            raise NotImplementedError("disable_self_attn is unused by the SDXL config")
        attn_cls = self.ATTENTION_MODES[attn_mode]
        self.add_lora = add_lora
        self.image_cross = image_cross
        self.rgb_predict = rgb_predict
        self.use_prev_weights_imp_sample = use_prev_weights_imp_sample
        self.imp_sample_next_step = imp_sample_next_step
        self.disable_self_attn = disable_self_attn
        self.rendered_feat = None
        self.choices: Optional[Sequence[int]] = None
        self.attn1 = attn_cls(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                              add_lora=add_lora, context_dim=None)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = attn_cls(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                              dropout=dropout, add_lora=add_lora)
        if image_cross:
            self.pose_emb_layers = Linear(2 * dim, dim, bias=False)
            nn.init.eye_(self.pose_emb_layers.weight)
            self.pose_featurenerf = NerfSDModule(
                mode=mode, out_channels=dim, far_plane=far, num_samples=num_samples,
                rgb_predict=rgb_predict, average=average, num_freqs=num_freqs, stratified=stratified,
                imp_sampling_percent=imp_sampling_percent, near_plane=near_plane)
            self.renderer = VolRender()
        self.norm1 = LayerNorm(dim)
        self.norm2 = LayerNorm(dim)
        self.norm3 = LayerNorm(dim)
        self.checkpoint = checkpoint
        self._ctxref_cache = None

    # ---- FeatureNeRF -------------------------------------------------------------------------
    def context_ref_tokens(self, batch: int) -> torch.Tensor:
        """Reference tokens per CFG row from the stored `references` buffer (sample.py:85-96):
        row group 0 -> the 'null' reference (last row) repeated n times, other groups -> the chosen
        real references.  bf16 [batch*n*hw, c]; cached (static across the sampling loop)."""
        live = self.__dict__.get("_live_ctxref")
        if live is not None:  # UNetModel.forward(input_ref=...): tokens of the live reference stream
            tok, n = live[:2]
            if len(live) > 2 and live[2] is not None:   # produced on the side stream (forward_train)
                torch.cuda.current_stream(tok.device).wait_event(live[2])
            assert tok.shape[0] % (batch * n) == 0
            self._ctxref_cache = (("live", tok.data_ptr()), tok, n)
            return tok
        refs = self.references
        choices = list(self.choices) if self.choices is not None else list(range(refs.shape[0] - 1))
        key = (batch, tuple(choices), refs.data_ptr())
        if self._ctxref_cache is not None and self._ctxref_cache[0] == key:
            return self._ctxref_cache[1]
        rows = 3 if batch % 3 == 0 else 2
        bs = batch // rows
        n = len(choices)
        real = refs[:-1][choices]                          # [n, hw, c]
        null = refs[-1:].expand(n, -1, -1)
        per_group = [null] + [real] * (rows - 1)
        full = torch.stack([g for g in per_group for _ in range(bs)])  # [batch, n, hw, c]
        tok = to_tokens(full.contiguous())
        self._ctxref_cache = (key, tok, n)
        return tok

    def reference_tokens(self, cams, xref_tok, n, kv, nkv, batch, hw, mask_ref=None):
        """reference_attn (reference :571-598) in token layout -> (rendered bf16 [batch*hw, c],
        fg [batch,hw], alphas [batch,hw,d], rgb [batch,hw,3]).  mask_ref: padding masks of the
        reference views [batch, n, 1, H, W] (None at inference); UNetModel.forward installs the
        call's masks as `_mask_ref` on every pose block."""
        nerf = self.pose_featurenerf
        d = nerf.raymarcher.num_samples
        if mask_ref is None:
            mask_ref = self.__dict__.get("_mask_ref")
        classes = self.__dict__.get("_row_classes")
        if classes is not None and mask_ref is None and classes[1].shape[0] == batch and classes[0].shape[0] < batch:
            # Rows with the same cameras AND the same reference tokens render the same features: the encoding
            # (points, G, hpre, gather / view softmax, W2, decoder) depends on nothing else.  sample.py passes
            # `pose * 3` and real references for guidance rows 1 and 2 (:85-96), a sweep batches several
            # prompts of ONE target camera: encode one representative row per class, expand by index.
            uniq, inverse = classes
            bu = uniq.shape[0]
            c_tok = xref_tok.shape[-1]
            xref_u = xref_tok.view(batch, n * hw, c_tok).index_select(0, uniq).reshape(bu * n * hw, c_tok)
            feats_u, raw_u, dists, _ = nerf.encode_tokens(cams.index_select(0, uniq).contiguous(), xref_u, bu, n, hw)
            feats = feats_u.view(bu, hw * d, -1).index_select(0, inverse).reshape(batch * hw * d, -1)
            raw = raw_u.view(bu, hw * d, -1).index_select(0, inverse).reshape(batch * hw * d, -1)
        else:
            feats, raw, dists, _ = nerf.encode_tokens(cams, xref_tok, batch, n, hw, mask_ref=mask_ref)
        # feats += attn2(norm2(feats), context): the block's own norm2 / attn2 over every sample
        fn = self.norm2.tokens(feats)
        feats = self.attn2.tokens(fn, batch, hw * d, kv=kv, nkv=nkv, residual=feats, out=feats)
        if raw.shape[-1] != 4:
            raw4 = torch.zeros(raw.shape[0], 4, device=raw.device, dtype=torch.float32)
            raw4[:, 3] = raw[:, -1]
            raw = raw4
        c = feats.shape[-1]
        return ops.nerf_volrender(feats, raw, dists, batch, hw, d, c)

    def reference_attn(self, x, context_ref, context, pose, prev_weights, mask_ref):
        """Reference-signature variant: context_ref [b, n, hw, c], context [b, 77, ctx]."""
        b, n, hw, c = context_ref.shape
        cams = pack_pose(pose, x.device)
        kv = self.attn2.project_context(to_tokens(context))
        rendered, fg, alphas, rgb = self.reference_tokens(cams, to_tokens(context_ref), n, kv,
                                                          context.shape[1], b, hw, mask_ref=mask_ref)
        d = alphas.shape[-1]
        return (ops.cast_f32(rendered).view(b, hw, c), fg.view(b, hw, 1), None,
                alphas.view(b, hw, d, 1), rgb if self.rgb_predict else None)

    # ---- token fast path -----------------------------------------------------------------------
    def tokens(self, x, batch, n, ctx_tok, nctx, cams=None, kv=None):
        """x bf16 [batch*n, c] (updated in place where possible) -> (x, aux | None).
        kv: this block's K|V projection of the context if the caller already computed it
        (UNetModel projects the context for ALL blocks in one GEMM per step)."""
        aux = None
        x = self.attn1.tokens(self.norm1.tokens(x), batch, n, residual=x, out=x)
        if kv is None:
            kv = self.attn2.project_context(ctx_tok)
        x = self.attn2.tokens(self.norm2.tokens(x), batch, n, kv=kv, nkv=nctx, residual=x, out=x)
        if self.image_cross and cams is not None:
            if self.rendered_feat is None:
                xref_tok = self.context_ref_tokens(batch)
                n_views = self._ctxref_cache[2]
                rendered, fg, alphas, rgb = self.reference_tokens(cams, xref_tok, n_views, kv, nctx,
                                                                  batch, n)
                # persistent buffer: a captured CUDA graph keeps reading this address after
                # clear_rendered_feat() and the next image's step 0
                buf = self.__dict__.get("_rendered_buf")
                if buf is None or buf.shape != rendered.shape or buf.device != rendered.device:
                    buf = rendered
                    self.__dict__["_rendered_buf"] = buf
                else:
                    buf.copy_(rendered)
                self.rendered_feat = buf
                aux = (fg, alphas, rgb)
            x = self.pose_emb_layers.tokens(x, a1=self.rendered_feat)  # Linear(cat[x, xref]) w/o the cat
        x = self.ff.tokens(self.norm3.tokens(x), residual=x, out=x)
        self._capture_out(x)
        return x, aux

    def _capture_out(self, x):
        """Reference-stream capture (UNetModel.capture_references): pose-capable blocks record the
        tokens they emit — what the reference's validation hook stores as `references`
        (diffusion.py:28-40, main.py:594-602).  x is updated in place downstream, hence the copy."""
        cap = self.__dict__.get("_capture")
        if cap is not None and self.image_cross:
            cap.append(x.clone())
            if self.__dict__.get("_capture_events") and x.is_cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(x.device))
                self.__dict__["_capture_ev"] = ev

    # ---- token fast path with the three LayerNorms folded into the GEMMs -------------------------
    def ln_packed(self):
        """gamma folded into the consuming weights (W' = W diag(gamma), bf16), beta into the bias
        (b' = b + W beta), plus the column sums of W' the epilogue needs: LN(x) W^T + b ==
        rstd (x W'^T - mu colsum(W')) + b'.  The statistics (mu, rstd) come from the (sum, sumsq)
        partials the PRODUCING GEMM's epilogue wrote, so no LayerNorm kernel and no normalised
        copy of x exist on this path."""
        dev = self.norm1.weight.device
        p = self.__dict__.get("_lnpk")
        if p is not None and p["dev"] == dev:
            return p
        f = lambda t: t.detach().float()

        def fold(w, b, norm):
            g, beta = f(norm.weight), f(norm.bias)
            wf = f(w)
            wq = (wf * g[None, :]).to(bf16).contiguous()
            bias = wf @ beta + (f(b) if b is not None else 0.0)
            return wq, bias.contiguous()

        a1, a2 = self.attn1, self.attn2
        wqkv, bqkv = fold(torch.cat([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], 0), None, self.norm1)
        wq2, bq2 = fold(a2.to_q.weight, None, self.norm2)
        proj = self.ff.net[0].proj
        wff, bff = fold(proj.weight, proj.bias, self.norm3)
        wff, bff = pack_geglu(wff, bff)  # interleave AFTER folding; column sums follow the packed rows
        p = dict(dev=dev, wqkv=wqkv, bqkv=bqkv, cqkv=wqkv.float().sum(1).contiguous(),
                 wq2=wq2, bq2=bq2, cq2=wq2.float().sum(1).contiguous(),
                 wff=wff, bff=bff, cff=wff.float().sum(1).contiguous())
        self.__dict__["_lnpk"] = p
        return p

    def tokens_fused(self, x, stats, batch, n, ctx_tok, nctx, cams=None, kv=None):
        """Like `tokens`, with stats = fp32 [batch*n, c/64, 2] row moments of x (from the GEMM that
        produced x).  Returns (x, stats_of_new_x, aux)."""
        aux = None
        lp = self.ln_packed()
        M, c = x.shape
        S = c // 64
        dev = x.device
        a1, a2 = self.attn1, self.attn2
        p1, p2 = a1.packed(), a2.packed()
        inner = a1.heads * a1.dim_head
        eps = self.norm1.eps
        new_stats = lambda: torch.empty((M, S, 2), device=dev, dtype=torch.float32)
        # self-attention: LN1 folded into the QKV projection
        qkv = ops.gemm(x, lp["wqkv"], bias=lp["bqkv"], ln_stats=stats, ln_colsum=lp["cqkv"], ln_eps=eps)
        a = ops.attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], batch, a1.heads,
                          n, n, ldq=3 * inner, ldk=3 * inner, ldv=3 * inner)
        st = new_stats()
        x = ops.gemm(a, p1["wo"], bias=p1["bo"], residual=x, out=x, stats_out=st)
        # text cross-attention: LN2 folded into the query projection
        if kv is None:
            kv = a2.project_context(ctx_tok)
        q = ops.gemm(x, lp["wq2"], bias=lp["bq2"], ln_stats=st, ln_colsum=lp["cq2"], ln_eps=eps)
        a = ops.attention(q, kv[:, :inner], kv[:, inner:2 * inner], batch, a2.heads, n, nctx,
                          ldq=inner, ldk=kv.stride(0), ldv=kv.stride(0))
        st = new_stats()
        x = ops.gemm(a, p2["wo"], bias=p2["bo"], residual=x, out=x, stats_out=st)
        if self.image_cross and cams is not None:
            if self.rendered_feat is None:
                xref_tok = self.context_ref_tokens(batch)
                n_views = self._ctxref_cache[2]
                rendered, fg, alphas, rgb = self.reference_tokens(cams, xref_tok, n_views, kv, nctx,
                                                                  batch, n)
                buf = self.__dict__.get("_rendered_buf")
                if buf is None or buf.shape != rendered.shape or buf.device != rendered.device:
                    buf = rendered
                    self.__dict__["_rendered_buf"] = buf
                else:
                    buf.copy_(rendered)
                self.rendered_feat = buf
                aux = (fg, alphas, rgb)
            st = new_stats()
            x = self.pose_emb_layers.tokens(x, a1=self.rendered_feat, stats_out=st)
        # feed-forward: LN3 folded into the GEGLU projection
        h = ops.gemm(x, lp["wff"], bias=lp["bff"], geglu=True, ln_stats=st, ln_colsum=lp["cff"],
                     ln_eps=eps)
        st = new_stats()
        x = self.ff.net[2].tokens(h, residual=x, out=x, stats_out=st)
        self._capture_out(x)
        return x, st, aux

    # ---- reference-signature entry point ---------------------------------------------------------
    def forward(self, x, context=None, context_ref=None, pose=None, mask_ref=None, prev_weights=None,
                additional_tokens=None, n_times_crossframe_attn_in_self=0):
        b, n, c = x.shape
        xt = to_tokens(x).clone()
        ctx_tok = to_tokens(context)
        cams = pack_pose(pose, x.device) if (pose is not None and self.image_cross) else None
        fg = alphas = rgb = None
        if cams is not None and context_ref is not None and not hasattr(self, "references"):
            # training-style call: explicit reference tokens [(b n), hw, c]
            xt = self.attn1.tokens(self.norm1.tokens(xt), b, n, residual=xt, out=xt)
            kv = self.attn2.project_context(ctx_tok)
            xt = self.attn2.tokens(self.norm2.tokens(xt), b, n, kv=kv, nkv=context.shape[1],
                                   residual=xt, out=xt)
            nv = context_ref.shape[0] // b
            rendered, fg, alphas, rgb = self.reference_tokens(cams, to_tokens(context_ref), nv, kv,
                                                              context.shape[1], b, n, mask_ref=mask_ref)
            xt = self.pose_emb_layers.tokens(xt, a1=rendered)
            xt = self.ff.tokens(self.norm3.tokens(xt), residual=xt, out=xt)
        else:
            self.__dict__["_mask_ref"] = mask_ref
            try:
                xt, aux = self.tokens(xt, b, n, ctx_tok, context.shape[1], cams)
            finally:
                self.__dict__.pop("_mask_ref", None)
            if aux is not None:
                fg, alphas, rgb = aux
        if fg is not None:
            d = alphas.shape[-1]
            fg, alphas = fg.view(b, n, 1), alphas.view(b, n, d, 1)
            rgb = rgb if self.rgb_predict else None
        return ops.cast_f32(xt).view(b, n, c), fg, None, alphas, rgb


class SpatialTransformer(nn.Module, _Packed):
    """GroupNorm -> proj_in -> depth x BasicTransformerBlock -> proj_out -> + x (reference :684-886)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None,
                 disable_self_attn=False, use_linear=False, attn_type="softmax", use_checkpoint=True,
                 sdp_backend=None, image_cross=True, rgb_predict=False, far=2, num_samples=32,
                 add_lora=False, mode="feature-nerf", average=False, num_freqs=16,
                 use_prev_weights_imp_sample=False, stratified=False, poscontrol_interval=4,
                 imp_sampling_percent=0.9, near_plane=0.0):
        super().__init__()
        if not use_linear:
            raise NotImplementedError("use_linear_in_transformer=False (1x1 conv projections) is not built")
        if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)
            assert all(cd == context_dim[0] for cd in context_dim), "need homogeneous context_dim"
            context_dim = context_dim[0]
        self.in_channels = in_channels
        inner = n_heads * d_head
        assert inner == in_channels
        self.norm = Normalize(in_channels)
        self.image_cross = image_cross
        self.poscontrol_interval = poscontrol_interval
        self.proj_in = Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(
                inner, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                disable_self_attn=disable_self_attn, attn_mode=attn_type, checkpoint=use_checkpoint,
                sdp_backend=sdp_backend, image_cross=image_cross and (d % poscontrol_interval == 0),
                far=far, num_samples=num_samples,
                add_lora=add_lora and image_cross and (d % poscontrol_interval == 0),
                rgb_predict=rgb_predict, mode=mode, average=average, num_freqs=num_freqs,
                use_prev_weights_imp_sample=use_prev_weights_imp_sample,
                imp_sample_next_step=False, stratified=stratified,
                imp_sampling_percent=imp_sampling_percent, near_plane=near_plane)
            for d in range(depth)])
        self.proj_out = Linear(inner, in_channels)
        nn.init.zeros_(self.proj_out.weight)  # zero_module (reference :795)
        nn.init.zeros_(self.proj_out.bias)
        self.use_linear = use_linear

    def _pack(self, dev):
        return dict(g=self.norm.weight.detach().float().contiguous(),
                    b=self.norm.bias.detach().float().contiguous())

    def tokens(self, x, batch, hw, ctx_tok, nctx, cams=None, aux_out=None, kv_all=None):
        """x bf16 [batch*hw, c] -> same shape (new tensor).  kv_all: optional [batch*nctx, sum 2c]
        context projection of every block of the network (blocks know their column slice)."""
        p = self.packed()
        xn = ops.groupnorm(x, p["g"], p["b"], batch, hw, eps=self.norm.eps, silu=False)
        fused = LN_FUSED and (self.in_channels % 64 == 0)
        stats = None
        if fused:  # proj_in's epilogue emits the row moments the first block's LayerNorm needs
            stats = torch.empty((xn.shape[0], self.in_channels // 64, 2), device=xn.device,
                                dtype=torch.float32)
            h = self.proj_in.tokens(xn, stats_out=stats)
        else:
            h = self.proj_in.tokens(xn)
        for i, block in enumerate(self.transformer_blocks):
            use_pose = self.image_cross and (i % self.poscontrol_interval == 0)
            kv = None
            sl = block.__dict__.get("_kv_slice")
            if kv_all is not None and sl is not None:
                kv = kv_all[:, sl[0]:sl[0] + sl[1]]
            if fused:
                h, stats, aux = block.tokens_fused(h, stats, batch, hw, ctx_tok, nctx,
                                                   cams if use_pose else None, kv=kv)
            else:
                h, aux = block.tokens(h, batch, hw, ctx_tok, nctx, cams if use_pose else None, kv=kv)
            if aux is not None and aux_out is not None:
                aux_out.append(aux)
        return self.proj_out.tokens(h, residual=x)

    def forward(self, x, xr=None, context=None, contextr=None, pose=None, mask_ref=None,
                prev_weights=None):
        """Reference contract: x [B, c, H, W] -> 6-tuple (x, xr, fg_masks, prev_weights, alphas, rgbs)."""
        if xr is not None:
            raise NotImplementedError("reference-image stream (training path) is a later row of SURVEY §8f")
        if isinstance(context, list):
            context = context[0]
        b, c, h, w = x.shape
        aux: list = []
        cams = pack_pose(pose, x.device) if pose is not None else None
        for blk in self.transformer_blocks:
            blk.__dict__["_mask_ref"] = mask_ref
        try:
            y = self.tokens(ops.nchw_to_nhwc_bf16(x.float().contiguous()), b, h * w, to_tokens(context),
                            context.shape[1], cams, aux)
        finally:
            for blk in self.transformer_blocks:
                blk.__dict__.pop("_mask_ref", None)
        out = ops.nhwc_to_nchw_f32(y, b, h * w, c).view(b, c, h, w)
        if aux:
            d = aux[0][1].shape[-1]
            fg = [a[0].view(b, h * w, 1) for a in aux]
            al = [a[1].view(b, h * w, d, 1) for a in aux]
            rgb = [a[2] for a in aux]
            return out, None, fg, None, al, rgb
        return out, None, None, None, None, None
