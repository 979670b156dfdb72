"""UNetModel / ResBlock / Upsample / Downsample / TimestepEmbedSequential of the reference
(sgm/modules/diffusionmodules/openaimodel.py) on the sm_100a kernels.

Same constructor kwargs (configs/train_co3d_concept.yaml:29-54), module tree and state-dict keys
as the reference, so `sd_xl_base_1.0.safetensors` and the delta checkpoints load by name.
`UNetModel.forward` keeps the reference's call contract
    net(x, timesteps=, context=, y=, **{pose, mask_ref, drop_im}) -> (eps, fg_masks, alphas, rgbs)
and runs the whole network in bf16 token layout ([B*H*W, C], i.e. NHWC): activations only change
layout at the first and last convolution.

Per forward (UNet batch B):
  * timestep embedding -> time_embed/label_emb MLPs and ALL ResBlock emb_layers in one small-M
    kernel each (the ResBlocks then receive their slice as the conv epilogue's per-image bias);
  * ResBlock = GN+SiLU -> implicit-GEMM conv3x3 (+bias +emb) -> GN+SiLU -> conv3x3 (+bias +skip),
    with decoder skip concatenations consumed as two K-segments / two GN sources (never copied);
  * SpatialTransformer / BasicTransformerBlock: see ..attention.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .... import ops
from ...._lib import ACT_SILU
from ...prepack import pack_conv3x3, pack_conv3x3_im2col
from ..attention import Linear, SpatialTransformer, _Packed, invalidate_all_packed, to_tokens
from ..utils_cameraray import pack_pose

bf16 = torch.bfloat16


# training step: run the no-grad reference stream on a side CUDA stream, concurrently with the taped
# main stream (CD360_REF_OVERLAP=0 serialises them again, for A/B measurements)
import os as _os
OVERLAP_REF_STREAM = _os.environ.get("CD360_REF_OVERLAP", "1") != "0"


class GroupNorm32(nn.GroupNorm):
    """Parameter holder; the arithmetic is cd360_groupnorm_silu_bf16 (fp32 statistics)."""


def normalization(channels):
    return GroupNorm32(32, channels)


class TimestepBlock(nn.Module):
    pass


class Upsample(nn.Module, _Packed):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1, third_up=False,
                 kernel_size=3, scale_factor=2):
        super().__init__()
        assert dims == 2 and use_conv and kernel_size == 3 and scale_factor == 2
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding)

    def _pack(self, dev):
        return dict(w=pack_conv3x3(self.conv.weight.detach()), b=self.conv.bias.detach().float().contiguous())

    def tokens(self, x, batch, h, w):
        p = self.packed()
        up = ops.upsample_nearest2x(x, batch, h, w)
        return ops.conv3x3(up, p["w"], batch, 2 * h, 2 * w, bias=p["b"])


class Downsample(nn.Module, _Packed):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1, third_down=False):
        super().__init__()
        assert dims == 2 and use_conv
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def _pack(self, dev):
        return dict(w=pack_conv3x3(self.op.weight.detach()), b=self.op.bias.detach().float().contiguous())

    def tokens(self, x, batch, h, w):
        p = self.packed()
        col = ops.im2col3x3_s2(x, batch, h, w)
        return ops.gemm(col, p["w"], bias=p["b"])


class ResBlock(TimestepBlock, _Packed):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False,
                 kernel_size=3, exchange_temb_dims=False, skip_t_emb=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv or skip_t_emb or dims != 2 or kernel_size != 3:
            raise NotImplementedError("ResBlock variant not used by the SDXL config")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_checkpoint = use_checkpoint
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        out_conv = nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)
        nn.init.zeros_(out_conv.weight)  # zero_module (reference :319-327)
        nn.init.zeros_(out_conv.bias)
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(),
                                        nn.Dropout(p=dropout), out_conv)
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def _pack(self, dev):
        f = lambda t: t.detach().float().contiguous()
        p = dict(g1=f(self.in_layers[0].weight), b1=f(self.in_layers[0].bias),
                 w1=pack_conv3x3(self.in_layers[2].weight.detach()), cb1=f(self.in_layers[2].bias),
                 g2=f(self.out_layers[0].weight), b2=f(self.out_layers[0].bias),
                 w2=pack_conv3x3(self.out_layers[3].weight.detach()), cb2=f(self.out_layers[3].bias))
        if not isinstance(self.skip_connection, nn.Identity):
            p["ws"] = self.skip_connection.weight.detach().reshape(self.out_channels, self.channels).to(bf16).contiguous()
            p["bs"] = f(self.skip_connection.bias)
        return p

    def tokens(self, x, batch, h, w, emb_out, skip=None):
        """x bf16 [batch*h*w, c0]; skip: optional second tensor [batch*h*w, c1] forming the channel
        concat [x | skip] (decoder).  emb_out fp32 [batch, out_channels] = emb_layers(emb)."""
        p = self.packed()
        hw = h * w
        hn = ops.groupnorm(x, p["g1"], p["b1"], batch, hw, x1=skip, eps=self.in_layers[0].eps, silu=True)
        h1 = ops.conv3x3(hn, p["w1"], batch, h, w, bias=p["cb1"], row_bias=emb_out)
        hn2 = ops.groupnorm(h1, p["g2"], p["b2"], batch, hw, eps=self.out_layers[0].eps, silu=True)
        if "ws" in p:
            xs = ops.gemm(x, p["ws"], bias=p["bs"], a1=skip)
        else:
            assert skip is None
            xs = x
        return ops.conv3x3(hn2, p["w2"], batch, h, w, bias=p["cb2"], residual=xs)

    def forward(self, x, emb):
        """Reference contract: x [B, C, H, W], emb [B, emb_channels] -> [B, Cout, H, W]."""
        b, c, h, w = x.shape
        lin = self.emb_layers[1]
        eo = ops.small_linear(emb.float().contiguous(), lin.weight.detach().to(bf16).contiguous(),
                              lin.bias.detach().float().contiguous(), act_in=ACT_SILU)
        y = self.tokens(ops.nchw_to_nhwc_bf16(x.float().contiguous()), b, h, w, eo)
        return ops.nhwc_to_nchw_f32(y, b, h * w, self.out_channels).view(b, self.out_channels, h, w)


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Container with the reference's name; UNetModel walks its children directly."""


class UNetModel(nn.Module, _Packed):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0.0, channel_mult=(1, 2, 4, 8), conv_resample=True,
                 dims=2, num_classes=None, use_checkpoint=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 transformer_depth=1, context_dim=None, disable_self_attentions=None,
                 num_attention_blocks=None, disable_middle_self_attn=False,
                 use_linear_in_transformer=False, spatial_transformer_attn_type="softmax",
                 adm_in_channels=None, use_fairscale_checkpoint=False, offload_to_cpu=False,
                 transformer_depth_middle=None,
                 # pose-conditioning arguments of the reference fork
                 image_cross_blocks=None, rgb=False, far=2.0, num_samples=32,
                 not_add_context_in_triplane=False, rgb_predict=False, add_lora=False,
                 mode="feature-nerf", average=False, num_freqs=16, use_prev_weights_imp_sample=False,
                 stratified=False, poscontrol_interval=4, imp_sampling_percent=0.9, near_plane=0.0):
        super().__init__()
        if dims != 2 or resblock_updown or use_scale_shift_norm or not conv_resample:
            raise NotImplementedError("UNet variant not used by the SDXL config")
        if num_classes != "sequential" or adm_in_channels is None:
            raise NotImplementedError("only num_classes='sequential' (SDXL vector conditioning) is built")
        if num_head_channels == -1:
            raise NotImplementedError("set num_head_channels (SDXL: 64)")
        if disable_self_attentions is not None or num_attention_blocks is not None:
            raise NotImplementedError("disable_self_attentions / num_attention_blocks are unused by the SDXL config")
        image_cross_blocks = list(image_cross_blocks or [])
        channel_mult = list(channel_mult)
        attention_resolutions = list(attention_resolutions)
        if isinstance(transformer_depth, int):
            transformer_depth = len(channel_mult) * [transformer_depth]
        transformer_depth = list(transformer_depth)
        if transformer_depth_middle is None:
            transformer_depth_middle = transformer_depth[-1]
        if isinstance(num_res_blocks, int):
            num_res_blocks = len(channel_mult) * [num_res_blocks]
        self.num_res_blocks = list(num_res_blocks)
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.rgb = rgb
        self.rgb_predict = rgb_predict
        self.attention_resolutions = attention_resolutions
        self.channel_mult = channel_mult
        self.num_classes = num_classes
        self.num_head_channels = num_head_channels
        self.use_checkpoint = use_checkpoint

        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.label_emb = nn.Sequential(nn.Sequential(nn.Linear(adm_in_channels, ted), nn.SiLU(),
                                                     nn.Linear(ted, ted)))

        def transformer(ch, depth, att_id):
            return SpatialTransformer(
                ch, ch // num_head_channels, num_head_channels, depth=depth, context_dim=context_dim,
                disable_self_attn=False, use_linear=use_linear_in_transformer,
                attn_type=spatial_transformer_attn_type, use_checkpoint=use_checkpoint,
                image_cross=(att_id in image_cross_blocks), rgb_predict=rgb_predict, far=far,
                num_samples=num_samples, add_lora=add_lora, mode=mode, average=average,
                num_freqs=num_freqs, use_prev_weights_imp_sample=use_prev_weights_imp_sample,
                stratified=stratified, poscontrol_interval=poscontrol_interval,
                imp_sampling_percent=imp_sampling_percent, near_plane=near_plane)

        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans = [model_channels]
        ch, ds, att_id = model_channels, 1, 0
        for level, mult in enumerate(channel_mult):
            for _ in range(self.num_res_blocks[level]):
                layers: List[nn.Module] = [ResBlock(ch, ted, dropout, out_channels=mult * model_channels,
                                                    use_checkpoint=use_checkpoint)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(transformer(ch, transformer_depth[level], att_id))
                    att_id += 1
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, True, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted, dropout, use_checkpoint=use_checkpoint),
            transformer(ch, transformer_depth_middle, att_id),
            ResBlock(ch, ted, dropout, use_checkpoint=use_checkpoint))
        att_id += 1
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(self.num_res_blocks[level] + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, dropout, out_channels=model_channels * mult,
                                   use_checkpoint=use_checkpoint)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(transformer(ch, transformer_depth[level], att_id))
                    att_id += 1
                if level and i == self.num_res_blocks[level]:
                    layers.append(Upsample(ch, True, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        out_conv = nn.Conv2d(model_channels, out_channels, 3, padding=1)
        nn.init.zeros_(out_conv.weight)  # zero_module (reference :971)
        nn.init.zeros_(out_conv.bias)
        self.out = nn.Sequential(normalization(ch), nn.SiLU(), out_conv)
        self.register_load_state_dict_post_hook(lambda module, incompatible: invalidate_all_packed(module))

    # ---- helpers the drivers reach for --------------------------------------------------------
    def pose_blocks(self):
        """(name, module) of every FeatureNeRF transformer block — the set
        DiffusionEngine.clear_rendered_feat / load_model_from_config walk (diffusion.py:165-169,
        sgm/util.py:231-235)."""
        for name, module in self.named_modules():
            parts = name.split(".")
            if len(parts) > 1 and parts[-2] == "transformer_blocks" and hasattr(module, "pose_emb_layers"):
                yield name, module

    def clear_rendered_feat(self):
        for _, m in self.pose_blocks():
            m.rendered_feat = None

    def set_reference_choices(self, choices: Optional[Sequence[int]]):
        """Which stored reference views condition sampling (sample.py:275-278's `choices`)."""
        for _, m in self.pose_blocks():
            m.choices = None if choices is None else list(choices)

    def register_references(self, refs: dict):
        """refs: {"<block name>.references" or "<block name>": tensor [R, hw, c]} (sgm/util.py:231-235)."""
        for name, m in self.pose_blocks():
            t = refs.get(name + ".references", refs.get(name))
            if t is not None:
                m.register_buffer("references", t)
                m._ctxref_cache = None

    @torch.no_grad()
    def capture_references(self, x_ref, timesteps, context, y):
        """The UNet's reference stream (openaimodel.py:79-111, attention.py:830-868): run the
        reference latents x_ref [R,4,L,L] through the SAME weights without pose conditioning and
        return {"<pose block name>": bf16 tokens [R, hw, c]} — the tensors the reference's validation
        hook saves as each pose block's `references` buffer (diffusion.py:28-40, main.py:594-602).
        Feed the result (plus the run's "null" row, sample.py:85-96) to `register_references`."""
        blocks = list(self.pose_blocks())
        for _, m in blocks:
            m.__dict__["_capture"] = []
        try:
            self.forward_tokens(x_ref, timesteps, context, y, pose=None)
        finally:
            caps = {name: m.__dict__.pop("_capture") for name, m in blocks}
        r = x_ref.shape[0]
        out = {}
        for name, lst in caps.items():
            assert len(lst) == 1, f"{name}: expected one reference-stream pass, got {len(lst)}"
            t = lst[0]
            out[name] = t.view(r, t.shape[0] // r, t.shape[1])
        return out

    def resblocks(self):
        return [m for m in self.modules() if isinstance(m, ResBlock)]

    def _pack(self, dev):
        f = lambda t: t.detach().float().contiguous()
        h = lambda t: t.detach().to(bf16).contiguous()
        rbs = self.resblocks()
        offs, o = [], 0
        for rb in rbs:
            offs.append(o)
            o += rb.out_channels
        p = dict(
            te0w=h(self.time_embed[0].weight), te0b=f(self.time_embed[0].bias),
            te2w=h(self.time_embed[2].weight), te2b=f(self.time_embed[2].bias),
            le0w=h(self.label_emb[0][0].weight), le0b=f(self.label_emb[0][0].bias),
            le2w=h(self.label_emb[0][2].weight), le2b=f(self.label_emb[0][2].bias),
            embw=torch.cat([h(rb.emb_layers[1].weight) for rb in rbs], 0).contiguous(),
            embb=torch.cat([f(rb.emb_layers[1].bias) for rb in rbs], 0).contiguous(),
            emb_off={id(rb): (off, rb.out_channels) for rb, off in zip(rbs, offs)},
            cin_w=pack_conv3x3_im2col(self.input_blocks[0][0].weight.detach(), 64),
            cin_b=f(self.input_blocks[0][0].bias),
            og=f(self.out[0].weight), ob=f(self.out[0].bias),
            cout_w=pack_conv3x3(self.out[2].weight.detach()), cout_b=f(self.out[2].bias))
        assert 9 * self.in_channels <= 64
        # context K|V projection weights of every transformer block, concatenated: one GEMM per
        # step instead of one small-M GEMM (M = 77*B rows) per block; the per-block packed copies
        # become views of this buffer
        from ..attention import BasicTransformerBlock
        blocks = [m for m in self.modules() if isinstance(m, BasicTransformerBlock)]
        if blocks:
            parts, off = [], 0
            for blk in blocks:
                wkv = blk.attn2.packed()["wkv"]
                blk.__dict__["_kv_slice"] = (off, wkv.shape[0])
                parts.append(wkv)
                off += wkv.shape[0]
            p["kvw"] = torch.cat(parts, 0).contiguous()
            for blk in blocks:
                o, n = blk.__dict__["_kv_slice"]
                blk.attn2.packed()["wkv"] = p["kvw"][o:o + n]
        return p

    # ---- forward -------------------------------------------------------------------------------
    def _run(self, layers, h, emb_all, p, batch, hh, ww, ctx_tok, nctx, cams, aux, skip=None):
        for layer in layers:
            if isinstance(layer, ResBlock):
                off, n = p["emb_off"][id(layer)]
                h = layer.tokens(h, batch, hh, ww, emb_all[:, off:off + n], skip=skip)
                skip = None
            elif isinstance(layer, SpatialTransformer):
                h = layer.tokens(h, batch, hh * ww, ctx_tok, nctx, cams, aux, kv_all=p.get("kv_all"))
            elif isinstance(layer, Downsample):
                h = layer.tokens(h, batch, hh, ww)
                hh, ww = hh // 2, ww // 2
            elif isinstance(layer, Upsample):
                h = layer.tokens(h, batch, hh, ww)
                hh, ww = hh * 2, ww * 2
            else:
                raise TypeError(type(layer))
        return h, hh, ww

    def forward(self, x, timesteps=None, context=None, y=None, timesteps2=None, **kwargs):
        """x [B,4,L,L] fp32, timesteps [B], context [B,77,ctx], y [B,adm] ->
        (eps [B,4,L,L] fp32, fg_mask_list, alphas_list, predicted_rgb_list)."""
        assert (y is not None), "must specify y: the model is class-conditional (num_classes='sequential')"
        mask_ref = kwargs.get("mask_ref")
        if mask_ref is not None:       # padding masks of the reference views (nerfsd_pytorch3d.py:61-70)
            for _, m in self.pose_blocks():
                m.__dict__["_mask_ref"] = mask_ref
            try:
                return self.forward(x, timesteps, context, y, timesteps2, **{k: v for k, v in kwargs.items() if k != "mask_ref"})
            finally:
                for _, m in self.pose_blocks():
                    m.__dict__.pop("_mask_ref", None)
        # (CPU tensors are rejected by the first kernel wrapper: there is no CPU path)
        xr = kwargs.get("input_ref")
        if xr is None:
            eps_tok, b, hh, ww, aux = self.forward_tokens(x, timesteps, context, y, kwargs.get("pose"),
                                                          kwargs.get("in_scale"))
        else:
            # The call shape of the training step, forward only (reference :1008-1051): the reference
            # latents input_ref [b, n, 4, L, L] form a second, pose-free stream through the same
            # weights (time embedding of `sigmas_ref` broadcast over the n views, text / vector
            # conditioning = the second halves of context / y) and every pose block of the main
            # stream reads that stream's tokens as its context_ref (attention.py:852-854).
            b, n = xr.shape[:2]
            assert context.shape[0] == b + b * n and y.shape[0] == b + b * n, \
                "context / y must hold the b target rows followed by the b*n reference rows"
            sig = kwargs.get("sigmas_ref", timesteps2)
            if sig is None:
                sig = torch.zeros_like(timesteps)
            t_ref = sig.reshape(b, 1).expand(b, n).reshape(b * n)
            caps = self.capture_references(xr.reshape(b * n, *xr.shape[2:]), t_ref, context[b:], y[b:])
            blocks = dict(self.pose_blocks())
            self.clear_rendered_feat()
            for name, m in blocks.items():
                t = caps[name]                                   # [b*n, hw, c]
                m.__dict__["_live_ctxref"] = (t.reshape(-1, t.shape[-1]), n)
                m._ctxref_cache = None
            try:
                eps_tok, b, hh, ww, aux = self.forward_tokens(x, timesteps, context[:b], y[:b],
                                                              kwargs.get("pose"), kwargs.get("in_scale"))
            finally:
                for m in blocks.values():
                    m.__dict__.pop("_live_ctxref", None)
                    m._ctxref_cache = None
                self.clear_rendered_feat()  # FeatureNeRF output is per call on this path (no caching)
        eps = ops.nhwc_to_nchw_f32(eps_tok, b, hh * ww, self.out_channels).view(b, self.out_channels, hh, ww)
        fg = [a[0].view(b, -1, 1) for a in aux]
        al = [a[1].view(b, a[1].shape[1], a[1].shape[2], 1) for a in aux]
        rgb = [a[2] for a in aux] if self.rgb_predict else []
        return eps, fg, al, rgb

    # ---- training step (explicit backward) ------------------------------------------
    def forward_train(self, x, timesteps=None, context=None, y=None, *, pose=None, input_ref=None,
                      sigmas_ref=None, in_scale=None, jitter=None, mask_ref=None, **_ignored):
        """The training-time call of the reference (openaimodel.py:1008-1093 with `input_ref`):
        the reference latents input_ref [b, n, 4, L, L] run as the no-grad reference stream
        (timesteps `sigmas_ref` broadcast over the views; second halves of context / y), the main
        stream runs taped.  Returns (eps fp32 tokens [b*L*L, 4], aux [(pose block, (fg [b,hw],
        alphas [b,hw,d], rgb [b,hw,3]))] in execution order, tape) for `backward`."""
        from .. import train_path
        b, n = input_ref.shape[:2]
        assert context.shape[0] == b + b * n and y.shape[0] == b + b * n, \
            "context / y must hold the b target rows followed by the b*n reference rows"
        sig = sigmas_ref if sigmas_ref is not None else torch.zeros_like(timesteps)
        t_ref = sig.reshape(b, 1).expand(b, n).reshape(b * n)
        blocks = dict(self.pose_blocks())
        # The reference stream is enqueued on a SIDE CUDA stream and the taped main stream on the
        # current one: both are chains of small launches (M = 256 ... 4096 rows) that leave most SMs
        # idle, so they overlap almost completely.  Each pose block of the main stream waits for the
        # event recorded when the reference stream emitted that block's tokens; the side stream is
        # joined before returning (a CUDA-graph capture records this as a fork / join).
        x_ref = input_ref.reshape(b * n, *input_ref.shape[2:])
        cur = side = None
        # (the first call after the packs were (re)built runs serialised: operand packs are created
        # lazily on whichever stream needs them first)
        if x.is_cuda and OVERLAP_REF_STREAM and self.__dict__.get("_packs_warm"):
            cur = torch.cuda.current_stream(x.device)
            side = self.__dict__.get("_side_stream")
            if side is None or side.device != x.device:
                side = torch.cuda.Stream(device=x.device)
                self.__dict__["_side_stream"] = side
            side.wait_stream(cur)
        for m in blocks.values():
            m.__dict__["_capture_events"] = side is not None
            m.__dict__["_mask_ref"] = mask_ref       # reference padding masks (nerfsd_pytorch3d.py:61-70)
        try:
            with torch.no_grad():
                if side is not None:
                    with torch.cuda.stream(side):
                        caps = self.capture_references(x_ref, t_ref, context[b:], y[b:])
                else:
                    caps = self.capture_references(x_ref, t_ref, context[b:], y[b:])
            for name, m in blocks.items():
                t = caps[name]   # stays referenced (tape / caps) until after the join below
                m.__dict__["_live_ctxref"] = (t.reshape(-1, t.shape[-1]), n, m.__dict__.pop("_capture_ev", None))
                m._ctxref_cache = None
            out = train_path.unet_forward(self, x, timesteps, context[:b], y[:b], pose, in_scale, jitter,
                                          cond_grad=bool(self.__dict__.get("cond_grad", True)))
            self.__dict__["_packs_warm"] = True
            return out
        finally:
            if side is not None:
                cur.wait_stream(side)
            for m in blocks.values():
                m.__dict__.pop("_live_ctxref", None)
                m.__dict__.pop("_capture_events", None)
                m.__dict__.pop("_capture_ev", None)
                m.__dict__.pop("_mask_ref", None)
                m._ctxref_cache = None

    def backward(self, tape, deps, daux_of=None):
        """Gradients of the pose weights (into `.grad`, fp32) from deps = dL/d(eps tokens) bf16
        [b*L*L, 64] (columns >= 4 zero) and daux_of = {id(pose block): (dfg, dalphas, drgb)}.
        Returns the conditioning gradients {"crossattn": fp32 [b, 77, ctx], "vector": fp32 [b, adm]} of
        the b TARGET rows (the reference-view rows get none: that stream is no_grad in the reference,
        openaimodel.py:95-108) unless `self.cond_grad` was set False before forward_train."""
        from .. import train_path
        return train_path.unet_backward(self, tape, deps, daux_of or {})

    def forward_tokens(self, x, timesteps, context, y, pose=None, in_scale=None, batch=None):
        """Same as forward but returns eps in token layout: fp32 [B*L*L, out_channels].
        in_scale: optional fp32 [B] multiplied into x on load (the denoiser's c_in).
        batch: UNet batch B when x holds only B / rows distinct latents (CFG rows replicate the
        latent: row b reads x[b % x.shape[0]]).  context may be pre-converted bf16 tokens
        [B*nctx, ctx] (then pass nctx via context.shape[0] // B) and pose a packed [B, n+1, 16]."""
        p = self.packed()
        src_b, cin, hh, ww = x.shape
        b = src_b if batch is None else batch
        assert y.shape[0] == b
        dev = x.device
        # embeddings
        t_emb = ops.timestep_embedding(timesteps.to(device=dev, dtype=torch.float32).contiguous(),
                                       self.model_channels)
        e1 = ops.small_linear(t_emb, p["te0w"], p["te0b"], act_out=ACT_SILU)
        emb = ops.small_linear(e1, p["te2w"], p["te2b"])
        l1 = ops.small_linear(y.float().contiguous(), p["le0w"], p["le0b"], act_out=ACT_SILU)
        emb = ops.small_linear(l1, p["le2w"], p["le2b"], add=emb)
        emb_all = ops.small_linear(emb, p["embw"], p["embb"], act_in=ACT_SILU)   # every ResBlock's emb_layers
        if context.dim() == 3:
            assert context.shape[0] == b
            ctx_tok, nctx = to_tokens(context), context.shape[1]
        else:
            ctx_tok, nctx = context, context.shape[0] // b
        cams = pack_pose(pose, dev) if pose is not None else None
        aux: list = []
        p["kv_all"] = ops.gemm(ctx_tok, p["kvw"]) if "kvw" in p else None
        # input conv: Cin=4 -> im2col (K=36 padded to 64) + GEMM
        col = ops.im2col3x3_nchw(x.float().contiguous(), 64, scale=in_scale, batch=b)
        h = ops.gemm(col, p["cin_w"], bias=p["cin_b"])
        hs = [h]
        for block in list(self.input_blocks)[1:]:
            h, hh, ww = self._run(block, h, emb_all, p, b, hh, ww, ctx_tok, nctx, cams, aux)
            hs.append(h)
        h, hh, ww = self._run(self.middle_block, h, emb_all, p, b, hh, ww, ctx_tok, nctx, cams, aux)
        for block in self.output_blocks:
            h, hh, ww = self._run(block, h, emb_all, p, b, hh, ww, ctx_tok, nctx, cams, aux, skip=hs.pop())
        hn = ops.groupnorm(h, p["og"], p["ob"], b, hh * ww, eps=self.out[0].eps, silu=True)
        eps = ops.conv3x3(hn, p["cout_w"], b, hh, ww, bias=p["cout_b"], out_fp32=True)
        return eps, b, hh, ww, aux
