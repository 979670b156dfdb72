"""VAE decoder of the first stage on the sm_100a kernels.

Reference: sgm/modules/diffusionmodules/model.py — `Decoder` (:604-757), `ResnetBlock` (:94-151),
`Upsample` (:58-71), `MemoryEfficientAttnBlock` / `AttnBlock` (:161-266), `Normalize` (:52-55).  Same
class names, constructor kwargs and state-dict keys (`conv_in`, `mid.block_1`, `mid.attn_1.{norm,q,k,
v,proj_out}`, `mid.block_2`, `up.<i>.block.<j>.{norm1,conv1,norm2,conv2,nin_shortcut}`,
`up.<i>.upsample.conv`, `norm_out`, `conv_out`), so `sdxl_vae.safetensors` loads by name.

Data path: bf16 NHWC tokens between kernels like the UNet — 3x3 convolutions are the tcgen05 implicit
GEMM (128-pixel row segments for images wider than 128), GroupNorm(32, eps 1e-6)+swish is
cd360_groupnorm_silu_bf16, nearest upsampling cd360_upsample_nearest2x_bf16.  The mid-block attention
is ONE head of width C (512) over all pixels — outside the 64-wide flash kernel — and runs once per
image, so it is two tcgen05 GEMMs around a row softmax: S = q k^T (fp32), P = softmax(S / sqrt(C))
(cd360_softmax_rows_f32_bf16), O = P v.  There is no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .... import ops
from ...prepack import pack_conv3x3, pack_conv3x3_im2col, pack_conv3x3_padded
from ..attention import _Packed

bf16 = torch.bfloat16


def _f(t):
    return t.detach().float().contiguous()


def Normalize(in_channels, num_groups=32):
    """Parameter holder (reference :52-55); the arithmetic is cd360_groupnorm_silu_bf16."""
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def nonlinearity(x):
    raise NotImplementedError("swish is fused into cd360_groupnorm_silu_bf16; there is no stand-alone module path")


class Upsample(nn.Module, _Packed):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("resamp_with_conv=False is unused by the shipped config")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def _pack(self, dev):
        return dict(w=pack_conv3x3(self.conv.weight.detach()), b=_f(self.conv.bias))

    def tokens(self, x, batch, h, w):
        p = self.packed()
        up = ops.upsample_nearest2x(x, batch, h, w)
        return ops.conv3x3(up, p["w"], batch, 2 * h, 2 * w, bias=p["b"])


class ResnetBlock(nn.Module, _Packed):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512):
        super().__init__()
        if temb_channels > 0 or conv_shortcut:
            raise NotImplementedError("the VAE decoder builds ResnetBlock with temb_channels=0 and a 1x1 shortcut")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def _pack(self, dev):
        p = dict(g1=_f(self.norm1.weight), b1=_f(self.norm1.bias), w1=pack_conv3x3(self.conv1.weight.detach()),
                 cb1=_f(self.conv1.bias), g2=_f(self.norm2.weight), b2=_f(self.norm2.bias),
                 w2=pack_conv3x3(self.conv2.weight.detach()), cb2=_f(self.conv2.bias))
        if self.in_channels != self.out_channels:
            p["ws"] = self.nin_shortcut.weight.detach().reshape(self.out_channels, self.in_channels).to(bf16).contiguous()
            p["bs"] = _f(self.nin_shortcut.bias)
        return p

    def tokens(self, x, batch, h, w):
        """x bf16 [batch*h*w, cin] -> [batch*h*w, cout] (reference :131-151, temb None)."""
        p = self.packed()
        hw = h * w
        hn = ops.groupnorm(x, p["g1"], p["b1"], batch, hw, eps=self.norm1.eps, silu=True)
        h1 = ops.conv3x3(hn, p["w1"], batch, h, w, bias=p["cb1"])
        del hn
        hn2 = ops.groupnorm(h1, p["g2"], p["b2"], batch, hw, eps=self.norm2.eps, silu=True)
        del h1
        xs = ops.gemm(x, p["ws"], bias=p["bs"]) if "ws" in p else x
        return ops.conv3x3(hn2, p["w2"], batch, h, w, bias=p["cb2"], residual=xs)


class MemoryEfficientAttnBlock(nn.Module, _Packed):
    """Single-head self-attention over all pixels (reference :204-266)."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.attention_op = None

    def _pack(self, dev):
        c = self.in_channels
        lin = lambda m: m.weight.detach().reshape(c, c).to(bf16).contiguous()
        return dict(g=_f(self.norm.weight), b=_f(self.norm.bias),
                    wq=lin(self.q), bq=_f(self.q.bias), wk=lin(self.k), bk=_f(self.k.bias),
                    wv=lin(self.v), bv=_f(self.v.bias), wo=lin(self.proj_out), bo=_f(self.proj_out.bias))

    def tokens(self, x, batch, h, w):
        p = self.packed()
        hw, c = h * w, self.in_channels
        hn = ops.groupnorm(x, p["g"], p["b"], batch, hw, eps=self.norm.eps, silu=False)
        q = ops.gemm(hn, p["wq"], bias=p["bq"])
        k = ops.gemm(hn, p["wk"], bias=p["bk"])
        v = ops.gemm(hn, p["wv"], bias=p["bv"])
        del hn
        o = torch.empty_like(q)
        scores = torch.empty((hw, hw), device=x.device, dtype=torch.float32)
        probs = torch.empty((hw, hw), device=x.device, dtype=bf16)
        for i in range(batch):   # one image at a time: the score matrix is hw x hw (1 GB fp32 at 128x128 latents)
            rows = slice(i * hw, (i + 1) * hw)
            ops.gemm(q[rows], k[rows], out=scores)
            ops.softmax_rows(scores, scale=1.0 / math.sqrt(c), out=probs)
            ops.gemm(probs, ops.transpose_to_bf16(v[rows], ld_out=hw), out=o[rows])
        return ops.gemm(o, p["wo"], bias=p["bo"], residual=x)


AttnBlock = MemoryEfficientAttnBlock   # "vanilla" and "vanilla-xformers" compute the same function


def make_attn(in_channels, attn_type="vanilla", attn_kwargs=None):
    if attn_type not in ("vanilla", "vanilla-xformers"):
        raise NotImplementedError(f"attn_type {attn_type!r}: the shipped VAE uses vanilla-xformers")
    return MemoryEfficientAttnBlock(in_channels)


class Decoder(nn.Module, _Packed):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False,
                 tanh_out=False, use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if give_pre_end or tanh_out or use_linear_attn:
            raise NotImplementedError("give_pre_end / tanh_out / linear attention are unused by the shipped config")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.out_ch = out_ch
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        if 9 * z_channels > 64:
            raise NotImplementedError("conv_in runs as im2col with K = 9*z_channels padded to 64")
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def get_last_layer(self, **kwargs):
        return self.conv_out.weight

    def _pack(self, dev):
        n_pad = (self.out_ch + 3) // 4 * 4       # fp32 rows of the last conv must be 16-byte multiples
        cb = torch.zeros(n_pad, device=dev)
        cb[: self.out_ch] = self.conv_out.bias.detach().float()
        return dict(cin_w=pack_conv3x3_im2col(self.conv_in.weight.detach(), 64), cin_b=_f(self.conv_in.bias),
                    og=_f(self.norm_out.weight), ob=_f(self.norm_out.bias), n_pad=n_pad,
                    cout_w=pack_conv3x3_padded(self.conv_out.weight.detach(), self.conv_out.in_channels, n_pad),
                    cout_b=cb)

    def forward(self, z, **kwargs):
        """z fp32 [B, z_channels, h, w] -> image fp32 [B, out_ch, 8h, 8w] (reference :715-757)."""
        p = self.packed()
        self.last_z_shape = z.shape
        b, _, hh, ww = z.shape
        col = ops.im2col3x3_nchw(z.float().contiguous(), 64)
        h = ops.gemm(col, p["cin_w"], bias=p["cin_b"])
        del col
        h = self.mid.block_1.tokens(h, b, hh, ww)
        h = self.mid.attn_1.tokens(h, b, hh, ww)
        h = self.mid.block_2.tokens(h, b, hh, ww)
        for i_level in reversed(range(self.num_resolutions)):
            up = self.up[i_level]
            for i_block in range(self.num_res_blocks + 1):
                h = up.block[i_block].tokens(h, b, hh, ww)
                if len(up.attn) > 0:
                    h = up.attn[i_block].tokens(h, b, hh, ww)
            if i_level != 0:
                h = up.upsample.tokens(h, b, hh, ww)
                hh, ww = 2 * hh, 2 * ww
        hn = ops.groupnorm(h, p["og"], p["ob"], b, hh * ww, eps=self.norm_out.eps, silu=True)
        del h
        img = ops.conv3x3(hn, p["cout_w"], b, hh, ww, bias=p["cout_b"], out_fp32=True)
        del hn
        out = ops.nhwc_to_nchw_f32(img, b, hh * ww, p["n_pad"]).view(b, p["n_pad"], hh, ww)
        return out[:, : self.out_ch]
