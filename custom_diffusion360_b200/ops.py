"""Torch-tensor front end of the C ABI.  Tensors supply device memory and the stream only;
every arithmetic op below is a hand-written sm_100a kernel in libcd360.so.

All wrappers require CUDA tensors and raise `Cd360Error` on any non-zero status — there is no
fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, Cd360Error, GemmArgs, check

bf16 = torch.bfloat16
f32 = torch.float32


class LaunchStats:
    """Count of libcd360 kernel launches issued from this process (every wrapper below obtains
    the stream exactly once per C-ABI call; GroupNorm is two launches)."""
    launches = 0
    hook = None  # optional callable(name:str, flops:float) -> context manager, set by bench.py


def _stream() -> int:
    LaunchStats.launches += 1
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise Cd360Error(f"{name}: expected a CUDA tensor (libcd360 has no CPU path)")
    if t.dtype != dtype:
        raise Cd360Error(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise Cd360Error(f"{name}: expected a contiguous tensor")


import math as _math
import os as _os

# Split-K for small-M GEMMs (training step: M = 256 tokens against 13-26 MB of weights; weight
# gradients with K ~ 10^5): on by default, CD360_SPLITK=0 disables it for A/B runs.
SPLITK = _os.environ.get("CD360_SPLITK", "1") != "0"
# below ~2.5 k of K the fixed cost of a launch dominates and the extra finish launch does not pay
GEMM_SMALL = _os.environ.get("CD360_GEMM_SMALL", "1") != "0"   # mirrors pick_config's small-problem rule
GEMM_SMALL_MAX = int(_os.environ.get("CD360_GEMM_SMALL_MAX", "0")) or 148   # (one wave of single CTAs ...
GEMM_SMALL_MAXK = int(_os.environ.get("CD360_GEMM_SMALL_MAXK", "0")) or (1 << 30)  # ... when K is below this)
SPLITK_MIN_KB = int(_os.environ.get("CD360_SPLITK_MIN_KB", "40"))
_splitk_ws: dict = {}


def _splitk_plan(M, N, K, out, residual, bias, plain_epilogue):
    """Number of K splits for cd360_gemm_bf16 (1 = none): only when the tile grid covers at most half
    of the GPU, every split keeps >= 4 k-blocks, and the epilogue is something
    cd360_splitk_finish can apply (bias, residual, conversion)."""
    if not SPLITK or not plain_epilogue or (N & 3):
        return 1
    if out is not None and ((out.stride(0) & 3) or (out.data_ptr() & 15)):
        return 1
    if residual is not None and ((residual.stride(0) & 3) or (residual.data_ptr() & 7)):
        return 1
    if bias is not None and (bias.data_ptr() & 15):
        return 1
    singles = _math.ceil(M / 128) * _math.ceil(N / 128)
    small = GEMM_SMALL and (singles <= 74 or (singles <= GEMM_SMALL_MAX and K < GEMM_SMALL_MAXK))
    pair = N > 128 and M > 128 and not small                               # pick_config of gemm_tcgen05.cu
    units = _math.ceil(M / 256) * _math.ceil(N / 256) if pair else singles
    max_units = 74 if pair else 148
    nkb = _math.ceil(K / 64)
    if units * 2 > max_units or nkb < SPLITK_MIN_KB:
        return 1
    s = min(nkb // 4, max_units // units, (16 << 20) // max(M * N, 1))
    return s if s >= 2 else 1


def _splitk_workspace(numel: int, device) -> torch.Tensor:
    """fp32 scratch for the partial tiles (fully overwritten by every split-K launch)."""
    key = (str(device), torch.cuda.current_stream().cuda_stream)   # one per stream: streams may run concurrently
    ws = _splitk_ws.get(key)
    if ws is None or ws.numel() < numel:
        ws = torch.empty(max(numel, 16 << 20), device=device, dtype=f32)
        _splitk_ws[key] = ws
    return ws


def gemm(a, w, *, bias=None, row_bias=None, rows_per_group=0, residual=None, out=None,
         out_fp32=False, act=ACT_NONE, geglu=False, a1=None, block_n=0, max_ctas=0,
         lda=None, lda1=None, k0=None, k1=None, M=None, ln_stats=None, ln_colsum=None, ln_eps=1e-5,
         stats_out=None, k_splits=None):
    """out = epilogue(A @ W^T).  a: bf16 [M, K0] (row stride `lda` if given), optional second
    K-segment a1 [M, K1]; w: bf16 [N, K0+K1].
    ln_stats / ln_colsum: LayerNorm of the A rows folded into the epilogue (w, bias pre-folded with
    gamma / beta, see include/cd360.h); stats_out: fp32 [M, N_out/64, 2] receives the per-slab
    (sum, sumsq) of the output rows for the next folded LayerNorm.
    k_splits: None = heuristic (`_splitk_plan`), 1 = off, n > 1 = split the K loop n ways (partials
    accumulated in an fp32 scratch, bias / residual / conversion by cd360_splitk_finish)."""
    lib = _lib.load()
    _req(w, bf16, "w")
    M = a.shape[0] if M is None else M
    k0 = a.shape[-1] if k0 is None else k0
    lda = (a.stride(0) if a.dim() == 2 else k0) if lda is None else lda
    if a1 is not None:
        k1 = a1.shape[-1] if k1 is None else k1
        lda1 = a1.stride(0) if lda1 is None else lda1
    else:
        k1, lda1 = 0, 0
    N = w.shape[0]
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=f32 if out_fp32 else bf16)
    plain = (row_bias is None and act == ACT_NONE and not geglu and ln_stats is None and stats_out is None
             and block_n == 0 and max_ctas == 0)
    if k_splits is None:
        k_splits = _splitk_plan(M, N, k0 + k1, out, residual, bias, plain)
    elif k_splits > 1 and not plain:
        raise Cd360Error("split-K GEMM supports only bias / residual epilogues")
    if k_splits > 1:
        slices = lib.cd360_splitk_slices(k0, k1, int(k_splits))
        stride = (M * N + 3) // 4 * 4
        ws = _splitk_workspace(slices * stride, a.device)
        args = GemmArgs(
            a0=_ptr(a), lda0=lda, k0=k0, a1=_ptr(a1), lda1=lda1, k1=k1, w=_ptr(w), bias=0, row_bias=0,
            rows_per_group=0, ld_row_bias=0, residual=0, ldr=0, out=_ptr(ws), ldo=N, out_fp32=1, M=M, N=N,
            conv=0, B=0, H=0, W=0, C=0, act=ACT_NONE, geglu=0, block_n=0, max_ctas=0, ln_stats=0, ln_slabs=0,
            ln_eps=0.0, ln_colsum=0, stats_out=0, k_splits=int(k_splits), split_stride=stride)
        _run("gemm", 2.0 * M * N * (k0 + k1),
             lambda: check(lib.cd360_gemm_bf16(C.byref(args), _stream()), "cd360_gemm_bf16(split-K)"))
        _run("splitk_finish", 0.0, lambda: check(lib.cd360_splitk_finish(
            _ptr(ws), N, stride, slices, _ptr(bias), _ptr(residual),
            residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0),
            int(out.dtype == f32), M, N, _stream()), "cd360_splitk_finish"))
        return out
    args = GemmArgs(
        a0=_ptr(a), lda0=lda, k0=k0, a1=_ptr(a1), lda1=lda1, k1=k1, w=_ptr(w),
        bias=_ptr(bias), row_bias=_ptr(row_bias), rows_per_group=rows_per_group,
        ld_row_bias=(row_bias.stride(0) if row_bias is not None else 0),
        residual=_ptr(residual), ldr=(residual.stride(0) if residual is not None else 0),
        out=_ptr(out), ldo=out.stride(0), out_fp32=int(out.dtype == f32), M=M, N=N,
        conv=0, B=0, H=0, W=0, C=0, act=act, geglu=int(geglu), block_n=block_n, max_ctas=max_ctas,
        ln_stats=_ptr(ln_stats), ln_slabs=(ln_stats.shape[1] if ln_stats is not None else 0),
        ln_eps=float(ln_eps), ln_colsum=_ptr(ln_colsum), stats_out=_ptr(stats_out), k_splits=0, split_stride=0)
    _run("gemm", 2.0 * M * N * (k0 + k1),
         lambda: check(lib.cd360_gemm_bf16(C.byref(args), _stream()), "cd360_gemm_bf16"))
    return out


def gemm_tn(a_t, w_t, *, bias=None, out=None, out_fp32=True, k_splits=None):
    """out [M, N] = a_t^T @ w_t for a_t bf16 [K, M] and w_t bf16 [K, N] (row strides from the views): the
    contraction runs over the ROWS of both operands, which the kernel consumes in place as MN-major
    tcgen05 operands (cd360_gemm_args.tn).  The weight gradients dW = dY^T X of the training step."""
    lib = _lib.load()
    for name, t in (("a_t", a_t), ("w_t", w_t)):
        if not t.is_cuda or t.dtype != bf16 or t.dim() != 2:
            raise Cd360Error(f"gemm_tn: {name} must be a 2-D bf16 CUDA tensor (libcd360 has no CPU path)")
    K, M = a_t.shape
    N = w_t.shape[1]
    if w_t.shape[0] != K:
        raise Cd360Error(f"gemm_tn: contraction lengths differ ({K} vs {w_t.shape[0]})")
    if a_t.stride(1) != 1 or w_t.stride(1) != 1:
        raise Cd360Error("gemm_tn: operands must have unit column stride")
    if out is None:
        out = torch.empty((M, N), device=a_t.device, dtype=f32 if out_fp32 else bf16)
    if k_splits is None:
        k_splits = _splitk_plan(M, N, K, out, None, bias, True)
    common = dict(a0=_ptr(a_t), lda0=a_t.stride(0), k0=K, a1=0, lda1=0, k1=0, w=_ptr(w_t), row_bias=0,
                  rows_per_group=0, ld_row_bias=0, residual=0, ldr=0, M=M, N=N, conv=0, B=0, H=0, W=0, C=0,
                  act=ACT_NONE, geglu=0, block_n=0, max_ctas=0, ln_stats=0, ln_slabs=0, ln_eps=0.0, ln_colsum=0,
                  stats_out=0, tn=1, ldw=w_t.stride(0))
    if k_splits > 1:
        slices = lib.cd360_splitk_slices(K, 0, int(k_splits))
        stride = (M * N + 3) // 4 * 4
        ws = _splitk_workspace(slices * stride, a_t.device)
        args = GemmArgs(bias=0, out=_ptr(ws), ldo=N, out_fp32=1, k_splits=int(k_splits), split_stride=stride, **common)
        _run("gemm", 2.0 * M * N * K,
             lambda: check(lib.cd360_gemm_bf16(C.byref(args), _stream()), "cd360_gemm_bf16(tn, split-K)"))
        _run("splitk_finish", 0.0, lambda: check(lib.cd360_splitk_finish(
            _ptr(ws), N, stride, slices, _ptr(bias), 0, 0, _ptr(out), out.stride(0),
            int(out.dtype == f32), M, N, _stream()), "cd360_splitk_finish"))
        return out
    args = GemmArgs(bias=_ptr(bias), out=_ptr(out), ldo=out.stride(0), out_fp32=int(out.dtype == f32),
                    k_splits=0, split_stride=0, **common)
    _run("gemm", 2.0 * M * N * K, lambda: check(lib.cd360_gemm_bf16(C.byref(args), _stream()), "cd360_gemm_bf16(tn)"))
    return out


def _run(name, flops, fn):
    """Launch through the optional profiling hook (bench.py brackets launches with CUDA events)."""
    h = LaunchStats.hook
    return fn() if h is None else h(name, flops, fn)


def _op(name):
    """Route a whole wrapper through the profiling hook (non-tensor-core kernels: no FLOP count)."""
    def deco(fn):
        def wrapped(*a, **kw):
            h = LaunchStats.hook
            return fn(*a, **kw) if h is None else h(name, 0.0, lambda: fn(*a, **kw))
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def conv3x3(x, w, B, H, W, *, bias=None, row_bias=None, residual=None, out=None, out_fp32=False,
            block_n=0, max_ctas=0):
    """Stride-1 pad-1 3x3 convolution as implicit GEMM.  x: bf16 [B*H*W, C] (NHWC);
    w: bf16 [N, 9*C] packed (ky, kx, c)."""
    lib = _lib.load()
    _req(x, bf16, "x")
    _req(w, bf16, "w")
    Cin = x.shape[-1]
    N = w.shape[0]
    M = B * H * W
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=f32 if out_fp32 else bf16)
    args = GemmArgs(
        a0=_ptr(x), lda0=Cin, k0=9 * Cin, a1=0, lda1=0, k1=0, w=_ptr(w), bias=_ptr(bias),
        row_bias=_ptr(row_bias), rows_per_group=H * W,
        ld_row_bias=(row_bias.stride(0) if row_bias is not None else 0), residual=_ptr(residual),
        ldr=(residual.stride(0) if residual is not None else 0), out=_ptr(out), ldo=out.stride(0),
        out_fp32=int(out.dtype == f32), M=M, N=N, conv=1, B=B, H=H, W=W, C=Cin, act=ACT_NONE,
        geglu=0, block_n=block_n, max_ctas=max_ctas)
    _run("gemm", 2.0 * M * N * 9 * Cin,
         lambda: check(lib.cd360_gemm_bf16(C.byref(args), _stream()), "cd360_gemm_bf16(conv)"))
    return out


def geglu_pack_block(n_total: int) -> int:
    r = _lib.load().cd360_geglu_pack_block(n_total)
    if r < 0:
        check(r, "cd360_geglu_pack_block")
    return r


def attention(q, k, v, batch, heads, nq, nkv, *, out=None, ldq=None, ldk=None, ldv=None):
    """q: [batch*nq, >=heads*64] view (row stride ldq), k/v: [batch*nkv, ...]; returns
    [batch*nq, heads*64] bf16."""
    lib = _lib.load()
    ldq = q.stride(0) if ldq is None else ldq
    ldk = k.stride(0) if ldk is None else ldk
    ldv = v.stride(0) if ldv is None else ldv
    if out is None:
        out = torch.empty((batch * nq, heads * 64), device=q.device, dtype=bf16)
    _run("attention", 4.0 * batch * heads * nq * nkv * 64,
         lambda: check(lib.cd360_attention_bf16(_ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(out),
                                                out.stride(0), batch, heads, nq, nkv, _stream()),
                       "cd360_attention_bf16"))
    return out


_gn_ws: dict = {}


def groupnorm_workspace(batch: int, hw: int, device) -> torch.Tensor:
    n = _lib.load().cd360_groupnorm_workspace_floats(batch, hw)
    key = (str(device), int(n), torch.cuda.current_stream().cuda_stream)   # one per stream (concurrent streams)
    ws = _gn_ws.get(key)
    if ws is None:
        ws = torch.empty(int(n), device=device, dtype=f32)
        _gn_ws[key] = ws
    return ws


@_op("groupnorm")
def groupnorm(x0, gamma, beta, batch, hw, *, x1=None, eps=1e-5, silu=True, out=None,
              workspace=None):
    """GroupNorm(32)+optional SiLU over NHWC bf16 [batch*hw, c0 (+c1)]."""
    lib = _lib.load()
    _req(x0, bf16, "x0")
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    if out is None:
        out = torch.empty((batch * hw, c0 + c1), device=x0.device, dtype=bf16)
    if workspace is None:
        workspace = groupnorm_workspace(batch, hw, x0.device)
    LaunchStats.launches += 1  # stats + apply
    check(lib.cd360_groupnorm_silu_bf16(_ptr(x0), c0, _ptr(x1), c1, _ptr(gamma), _ptr(beta),
                                        _ptr(out), _ptr(workspace), batch, hw, eps, int(silu),
                                        _stream()), "cd360_groupnorm_silu_bf16")
    return out


@_op("layernorm")
def layernorm(x, gamma, beta, *, eps=1e-5, out=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib.cd360_layernorm_bf16(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), rows, c, eps,
                                   _stream()), "cd360_layernorm_bf16")
    return out


@_op("small_linear")
def small_linear(x, w, bias=None, *, add=None, act_in=ACT_NONE, act_out=ACT_NONE, out=None):
    lib = _lib.load()
    _req(x, f32, "x")
    _req(w, bf16, "w")
    batch, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((batch, n), device=x.device, dtype=f32)
    check(lib.cd360_small_linear(_ptr(x), _ptr(w), _ptr(bias), _ptr(add), _ptr(out), batch, n, k,
                                 act_in, act_out, _stream()), "cd360_small_linear")
    return out


@_op("other")
def timestep_embedding(t, dim, *, out=None):
    lib = _lib.load()
    _req(t, f32, "t")
    batch = t.shape[0]
    if out is None:
        out = torch.empty((batch, dim), device=t.device, dtype=f32)
    check(lib.cd360_timestep_embedding(_ptr(t), _ptr(out), batch, dim, _stream()),
          "cd360_timestep_embedding")
    return out


@_op("other")
def im2col3x3_nchw(x, kpad, *, scale=None, out=None, batch=None):
    """batch > x.shape[0] replicates the images cyclically (CFG rows) while loading."""
    lib = _lib.load()
    _req(x, f32, "x")
    src_b, cin, h, w = x.shape
    b = src_b if batch is None else batch
    if out is None:
        out = torch.empty((b * h * w, kpad), device=x.device, dtype=bf16)
    check(lib.cd360_im2col3x3_nchw_f32(_ptr(x), _ptr(scale), _ptr(out), b, src_b, cin, h, w, kpad,
                                       _stream()), "cd360_im2col3x3_nchw_f32")
    return out


@_op("im2col_s2")
def im2col3x3_s2(x, batch, h, w, *, out=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    c = x.shape[-1]
    if out is None:
        out = torch.empty((batch * (h // 2) * (w // 2), 9 * c), device=x.device, dtype=bf16)
    check(lib.cd360_im2col3x3_s2_bf16(_ptr(x), _ptr(out), batch, h, w, c, _stream()),
          "cd360_im2col3x3_s2_bf16")
    return out


@_op("upsample")
def upsample_nearest2x(x, batch, h, w, *, out=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    c = x.shape[-1]
    if out is None:
        out = torch.empty((batch * 4 * h * w, c), device=x.device, dtype=bf16)
    check(lib.cd360_upsample_nearest2x_bf16(_ptr(x), _ptr(out), batch, h, w, c, _stream()),
          "cd360_upsample_nearest2x_bf16")
    return out


def cfg_euler_step(x, eps, n_img, guidance_rows, hw, sigma_q, sigma, sigma_next, scale, scale_im,
                   *, denoised_out=None):
    lib = _lib.load()
    _req(x, f32, "x")
    _req(eps, f32, "eps")
    check(lib.cd360_cfg_euler_step(_ptr(x), _ptr(eps), _ptr(denoised_out), n_img, guidance_rows,
                                   hw, float(sigma_q), float(sigma), float(sigma_next),
                                   float(scale), float(scale_im), _stream()),
          "cd360_cfg_euler_step")
    return x


@_op("cfg_euler")
def cfg_euler_step_dev(x, eps, n_img, guidance_rows, hw, sigmas3, scale, scale_im, *,
                       denoised_out=None):
    lib = _lib.load()
    _req(x, f32, "x")
    _req(eps, f32, "eps")
    _req(sigmas3, f32, "sigmas3")
    check(lib.cd360_cfg_euler_step_dev(_ptr(x), _ptr(eps), _ptr(denoised_out), n_img, guidance_rows,
                                       hw, _ptr(sigmas3), float(scale), float(scale_im), _stream()),
          "cd360_cfg_euler_step_dev")
    return x


@_op("cast")
def cast_bf16(x, *, out=None):
    lib = _lib.load()
    _req(x, f32, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=bf16)
    check(lib.cd360_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()),
          "cd360_cast_f32_to_bf16")
    return out


@_op("cast")
def cast_f32(x, *, out=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=f32)
    check(lib.cd360_cast_bf16_to_f32(_ptr(x), _ptr(out), x.numel(), _stream()),
          "cd360_cast_bf16_to_f32")
    return out


@_op("other")
def nhwc_to_nchw_f32(x, batch, hw, c, *, out=None):
    lib = _lib.load()
    if out is None:
        out = torch.empty((batch, c, hw), device=x.device, dtype=f32)
    check(lib.cd360_nhwc_to_nchw_f32(_ptr(x), int(x.dtype == f32), _ptr(out), batch, hw, c,
                                     _stream()), "cd360_nhwc_to_nchw_f32")
    return out


@_op("other")
def nchw_to_nhwc_bf16(x, *, out=None):
    """fp32 [B, C, ...spatial] -> bf16 tokens [B*hw, C]."""
    lib = _lib.load()
    _req(x, f32, "x")
    b, c = x.shape[:2]
    hw = x.numel() // (b * c)
    if out is None:
        out = torch.empty((b * hw, c), device=x.device, dtype=bf16)
    check(lib.cd360_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), b, hw, c, _stream()),
          "cd360_nchw_f32_to_nhwc_bf16")
    return out


@_op("other")
def pointwise_conv_nchw(x, w, bias=None, *, scale=1.0, out=None):
    """fp32 NCHW 1x1 convolution with <= 8 channels either side (the VAE's post_quant_conv with
    1/scale_factor folded in): out = bias + scale * w @ x."""
    lib = _lib.load()
    _req(x, f32, "x")
    _req(w, f32, "w")
    b, cin = x.shape[:2]
    cout = w.shape[0]
    hw = x.numel() // (b * cin)
    if out is None:
        out = torch.empty((b, cout) + tuple(x.shape[2:]), device=x.device, dtype=f32)
    check(lib.cd360_pointwise_conv_nchw_f32(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), b, cin, cout, hw,
                                            float(scale), _stream()), "cd360_pointwise_conv_nchw_f32")
    return out


@_op("softmax")
def softmax_rows(s, *, scale=1.0, out=None):
    """softmax(scale * s) along the last dim: fp32 [rows, n] -> bf16 [rows, n]."""
    lib = _lib.load()
    _req(s, f32, "s")
    rows, n = s.shape
    if out is None:
        out = torch.empty((rows, n), device=s.device, dtype=bf16)
    check(lib.cd360_softmax_rows_f32_bf16(_ptr(s), s.stride(0), _ptr(out), out.stride(0), rows, n,
                                          float(scale), _stream()), "cd360_softmax_rows_f32_bf16")
    return out


@_op("nerf")
def nerf_points(cams, xy, depths, w_nv_geo, b_nv, b, n, res, d, kpe):
    lib = _lib.load()
    dev = cams.device
    hw = res * res
    pe = torch.empty((b * n * hw * d, kpe), device=dev, dtype=bf16)
    gidx = torch.empty((b * n * hw * d, 4), device=dev, dtype=torch.int32)
    gwgt = torch.empty((b * n * hw * d, 4), device=dev, dtype=f32)
    vlogit = torch.empty((b, n, hw * d), device=dev, dtype=f32)
    check(lib.cd360_nerf_points(_ptr(cams), _ptr(xy), _ptr(depths), _ptr(w_nv_geo), _ptr(b_nv),
                                _ptr(pe), _ptr(gidx), _ptr(gwgt), _ptr(vlogit), b, n, res, d, kpe,
                                _stream()), "cd360_nerf_points")
    return pe, gidx, gwgt, vlogit


@_op("nerf")
def nerf_combine(g, hpre, gidx, gwgt, vlogit, b, n, hw, d, c):
    lib = _lib.load()
    s = torch.empty((b * hw * d, c), device=g.device, dtype=bf16)
    vs = torch.empty((b, n, hw * d), device=g.device, dtype=f32)
    check(lib.cd360_nerf_combine(_ptr(g), g.stride(0), _ptr(hpre), _ptr(gidx), _ptr(gwgt),
                                 _ptr(vlogit), _ptr(s), _ptr(vs), b, n, hw, d, c, _stream()),
          "cd360_nerf_combine")
    return s, vs


@_op("nerf")
def nerf_mask_ref(xref_tok, mask_ref, bn, res, *, out=None):
    """xref_tok bf16 [bn*res*res, c] scaled row-wise by the nearest-resized padding masks
    mask_ref fp32 [bn, (1,) mh, mw] (nerfsd_pytorch3d.py:61-70)."""
    lib = _lib.load()
    _req(xref_tok, bf16, "xref_tok")
    m = mask_ref.reshape(bn, mask_ref.shape[-2], mask_ref.shape[-1])
    _req(m, f32, "mask_ref")
    if out is None:
        out = torch.empty_like(xref_tok)
    check(lib.cd360_nerf_mask_ref(_ptr(xref_tok), _ptr(m), _ptr(out), bn, res, m.shape[-2], m.shape[-1],
                                  xref_tok.shape[-1], _stream()), "cd360_nerf_mask_ref")
    return out


@_op("nerf")
def nerf_volrender(feats, raw, dists, b, hw, d, c):
    lib = _lib.load()
    dev = feats.device
    rendered = torch.empty((b * hw, c), device=dev, dtype=bf16)
    fg = torch.empty((b, hw), device=dev, dtype=f32)
    alphas = torch.empty((b, hw, d), device=dev, dtype=f32)
    rgb = torch.empty((b, hw, 3), device=dev, dtype=f32)
    check(lib.cd360_nerf_volrender(_ptr(feats), _ptr(raw), _ptr(dists), _ptr(rendered), _ptr(fg),
                                   _ptr(alphas), _ptr(rgb), b, hw, d, c, _stream()),
          "cd360_nerf_volrender")
    return rendered, fg, alphas, rgb


# ------------------------------------------------------------------------------------------------
# text conditioner (csrc/conditioner.cu)
# ------------------------------------------------------------------------------------------------
@_op("conditioner")
def embed_tokens(ids, tok_emb, pos_emb):
    """ids int32 [B, ctx] (device); tok_emb fp32 [V, w]; pos_emb fp32 [ctx, w] -> bf16 [B*ctx, w]."""
    lib = _lib.load()
    _req(ids, torch.int32, "ids")
    _req(tok_emb, f32, "tok_emb")
    _req(pos_emb, f32, "pos_emb")
    b, ctx = ids.shape
    vocab, w = tok_emb.shape
    if pos_emb.shape[0] < ctx:
        raise Cd360Error(f"embed_tokens: {ctx} tokens per sequence but only {pos_emb.shape[0]} positions")
    out = torch.empty((b * ctx, w), device=ids.device, dtype=bf16)
    check(lib.cd360_embed_tokens(_ptr(ids), _ptr(tok_emb), _ptr(pos_emb), _ptr(out), b * ctx, ctx, w, vocab,
                                 _stream()), "cd360_embed_tokens")
    return out


@_op("conditioner")
def attention_causal(q, k, v, batch, heads, n, *, out=None):
    """Causal softmax(QK^T/8)V, head dim 64; q/k/v 2-D bf16 views (row strides from the views)."""
    lib = _lib.load()
    if out is None:
        out = torch.empty((batch * n, heads * 64), device=q.device, dtype=bf16)
    check(lib.cd360_attention_causal_bf16(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0),
                                          _ptr(out), out.stride(0), batch, heads, n, _stream()),
          "cd360_attention_causal_bf16")
    return out


@_op("conditioner")
def gather_rows(x, idx):
    """x bf16 [R, c] (row stride from the view), idx int32 [B] -> fp32 [B, c]."""
    lib = _lib.load()
    _req(idx, torch.int32, "idx")
    if x.dtype != bf16 or not x.is_cuda:
        raise Cd360Error("gather_rows: expected a CUDA bf16 tensor")
    out = torch.empty((idx.shape[0], x.shape[1]), device=x.device, dtype=f32)
    check(lib.cd360_gather_rows_bf16_f32(_ptr(x), x.stride(0), _ptr(idx), _ptr(out), idx.shape[0], x.shape[1],
                                         x.shape[0], _stream()), "cd360_gather_rows_bf16_f32")
    return out


# ------------------------------------------------------------------------------------------------
# training step: backward / loss / optimiser kernels (csrc/train.cu, csrc/attention_bwd.cu)
# ------------------------------------------------------------------------------------------------
@_op("attention_bwd")
def attention_bwd(q, k, v, o, dout, batch, heads, nq, nkv, *, dq, dk=None, dv=None):
    """Gradients of `attention`.  q/k/v/o/dout and dq/dk/dv are 2-D views (row strides taken from
    the views, so slices of a fused QKV buffer work).  dk/dv None: query gradient only.
    Few keys x very many queries (reference_attn over the ray samples): dK / dV come from the
    query-split kernel (partial sums over `nsplit` CTAs per key tile meet in an fp32 scratch)."""
    lib = _lib.load()
    dev = q.device
    lse = torch.empty((batch, heads, nq), device=dev, dtype=f32)
    dsum = torch.empty((batch, heads, nq), device=dev, dtype=f32)
    key_tiles, q_tiles = (nkv + 127) // 128, (nq + 127) // 128      # 128 x 128 tiles (tcgen05 kernels)
    base_ctas = key_tiles * heads * batch
    split = dk is not None and base_ctas < 148 and q_tiles >= 16
    LaunchStats.launches += 1 if dk is None else (3 if split else 2)  # stats + dq (+ dkdv (+ finish))
    check(lib.cd360_attention_bwd_bf16(
        _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(o), o.stride(0),
        _ptr(dout), dout.stride(0), _ptr(dq), dq.stride(0), _ptr(None if split else dk),
        0 if dk is None or split else dk.stride(0), _ptr(None if split else dv),
        0 if dv is None or split else dv.stride(0), _ptr(lse), _ptr(dsum), batch, heads, nq, nkv,
        _stream()), "cd360_attention_bwd_bf16")
    if split:
        nsplit = max(1, min(q_tiles // 2, (2 * 148) // base_ctas))
        acc = torch.zeros((2, batch * nkv, heads * 64), device=dev, dtype=f32)
        check(lib.cd360_attention_bwd_kv_split_bf16(
            _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(dout), dout.stride(0),
            _ptr(lse), _ptr(dsum), _ptr(acc), _ptr(dk), dk.stride(0), _ptr(dv), dv.stride(0),
            batch, heads, nq, nkv, nsplit, _stream()), "cd360_attention_bwd_kv_split_bf16")
    return dq, dk, dv


@_op("layernorm_bwd")
def layernorm_bwd(x, gamma, dy, *, add=None, eps=1e-5, out=None):
    lib = _lib.load()
    _req(x, bf16, "x")
    _req(dy, bf16, "dy")
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib.cd360_layernorm_bwd_bf16(_ptr(x), _ptr(gamma), _ptr(dy), _ptr(add), _ptr(out), rows, c,
                                       eps, _stream()), "cd360_layernorm_bwd_bf16")
    return out


@_op("groupnorm_bwd")
def groupnorm_bwd(x0, gamma, beta, dy, batch, hw, *, x1=None, add0=None, add1=None, eps=1e-5,
                  silu=True):
    """Returns (dx0, dx1 | None).  add0 / add1 may be column slices (row stride from the view)."""
    lib = _lib.load()
    _req(x0, bf16, "x0")
    _req(dy, bf16, "dy")
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else x1.shape[-1]
    dev = x0.device
    dx0 = torch.empty((batch * hw, c0), device=dev, dtype=bf16)
    dx1 = None if x1 is None else torch.empty((batch * hw, c1), device=dev, dtype=bf16)
    ws = torch.empty(int(lib.cd360_groupnorm_bwd_workspace_floats(batch)), device=dev, dtype=f32)
    LaunchStats.launches += 1
    check(lib.cd360_groupnorm_silu_bwd_bf16(
        _ptr(x0), c0, _ptr(x1), c1, _ptr(gamma), _ptr(beta), _ptr(dy), _ptr(add0),
        0 if add0 is None else add0.stride(0), _ptr(add1), 0 if add1 is None else add1.stride(0),
        _ptr(dx0), _ptr(dx1), _ptr(ws), batch, hw, eps, int(silu), _stream()),
        "cd360_groupnorm_silu_bwd_bf16")
    return dx0, dx1


@_op("geglu_bwd")
def geglu_bwd(raw, dh, block=None):
    lib = _lib.load()
    _req(raw, bf16, "raw")
    _req(dh, bf16, "dh")
    rows, f = dh.shape
    out = torch.empty_like(raw)
    check(lib.cd360_geglu_bwd_bf16(_ptr(raw), _ptr(dh), _ptr(out), rows, f, f if block is None else block,
                                   _stream()), "cd360_geglu_bwd_bf16")
    return out


@_op("geglu_fwd")
def geglu_fwd(raw, block=None):
    """raw bf16 [rows, 2f] ([a | gate] in blocks of `block` columns) -> bf16 [rows, f] = a * gelu(gate)."""
    lib = _lib.load()
    _req(raw, bf16, "raw")
    rows, f2 = raw.shape
    f = f2 // 2
    out = torch.empty((rows, f), device=raw.device, dtype=bf16)
    check(lib.cd360_geglu_fwd_bf16(_ptr(raw), _ptr(out), rows, f, f if block is None else block, _stream()),
          "cd360_geglu_fwd_bf16")
    return out


@_op("add")
def add_bf16(a, b, *, out=None):
    lib = _lib.load()
    _req(a, bf16, "a")
    _req(b, bf16, "b")
    if out is None:
        out = torch.empty_like(a)
    check(lib.cd360_add_bf16(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "cd360_add_bf16")
    return out


@_op("other")
def silu_bwd(pre, dy, *, out=None):
    """out = dy * silu'(pre), fp32 (same shape)."""
    lib = _lib.load()
    _req(pre, f32, "pre")
    _req(dy, f32, "dy")
    if out is None:
        out = torch.empty_like(pre)
    check(lib.cd360_silu_bwd_f32(_ptr(pre), _ptr(dy), _ptr(out), pre.numel(), _stream()), "cd360_silu_bwd_f32")
    return out


@_op("transpose")
def transpose_to_bf16(x, *, ld_out=None):
    """x [rows, cols] bf16 / fp32 (row stride from the view) -> bf16 [cols, ld_out >= rows]."""
    lib = _lib.load()
    rows, cols = x.shape
    if ld_out is None:
        ld_out = (rows + 7) // 8 * 8
    out = torch.empty((cols, ld_out), device=x.device, dtype=bf16)
    check(lib.cd360_transpose_to_bf16(_ptr(x), int(x.dtype == f32), x.stride(0), _ptr(out), ld_out, rows,
                                      cols, _stream()), "cd360_transpose_to_bf16")
    return out


@_op("colsum")
def colsum(x, *, out=None):
    """fp32 [c] += column sums of bf16 x [rows, c]; `out` is accumulated into (zeros if None)."""
    lib = _lib.load()
    rows, c = x.shape
    if out is None:
        out = torch.zeros(c, device=x.device, dtype=f32)
    check(lib.cd360_colsum_bf16(_ptr(x), x.stride(0), _ptr(out), rows, c, _stream()), "cd360_colsum_bf16")
    return out


@_op("col2im_s2")
def col2im3x3_s2(dcol, batch, h, w, c):
    lib = _lib.load()
    _req(dcol, bf16, "dcol")
    out = torch.empty((batch * h * w, c), device=dcol.device, dtype=bf16)
    check(lib.cd360_col2im3x3_s2_bf16(_ptr(dcol), _ptr(out), batch, h, w, c, _stream()),
          "cd360_col2im3x3_s2_bf16")
    return out


@_op("upsample_bwd")
def upsample_nearest2x_bwd(g, batch, h, w):
    """g bf16 [batch*2h*2w, c] -> [batch*h*w, c]."""
    lib = _lib.load()
    _req(g, bf16, "g")
    c = g.shape[-1]
    out = torch.empty((batch * h * w, c), device=g.device, dtype=bf16)
    check(lib.cd360_upsample_nearest2x_bwd_bf16(_ptr(g), _ptr(out), batch, h, w, c, _stream()),
          "cd360_upsample_nearest2x_bwd_bf16")
    return out


@_op("nerf_bwd")
def nerf_volrender_bwd(feats, raw, dists, d_rendered, dfg, dalphas, drgb, b, hw, d, c):
    lib = _lib.load()
    dev = feats.device
    dfeats = torch.empty((b * hw * d, c), device=dev, dtype=bf16)
    draw = torch.empty((b * hw * d, 8), device=dev, dtype=bf16)
    check(lib.cd360_nerf_volrender_bwd(_ptr(feats), _ptr(raw), _ptr(dists), _ptr(d_rendered), _ptr(dfg),
                                       _ptr(dalphas), _ptr(drgb), _ptr(dfeats), _ptr(draw), b, hw, d, c,
                                       _stream()), "cd360_nerf_volrender_bwd")
    return dfeats, draw


@_op("nerf_bwd")
def nerf_combine_bwd(g, hpre, gidx, gwgt, vlogit, ds, b, n, hw, d, c):
    lib = _lib.load()
    dev = g.device
    dhpre = torch.empty((b * n * hw * d, c), device=dev, dtype=bf16)
    dlogit = torch.empty((b, n, hw * d), device=dev, dtype=f32)
    dg = torch.zeros((b * n * hw, g.stride(0)), device=dev, dtype=f32)
    check(lib.cd360_nerf_combine_bwd(_ptr(g), g.stride(0), _ptr(hpre), _ptr(gidx), _ptr(gwgt),
                                     _ptr(vlogit), _ptr(ds), _ptr(dhpre), _ptr(dlogit), _ptr(dg), b, n,
                                     hw, d, c, _stream()), "cd360_nerf_combine_bwd")
    return dhpre, dlogit, dg


@_op("nerf_bwd")
def nerf_nviews_geo_bwd(cams, dlogit, b, n):
    lib = _lib.load()
    dw = torch.zeros(198, device=cams.device, dtype=f32)
    check(lib.cd360_nerf_nviews_geo_bwd(_ptr(cams), _ptr(dlogit), _ptr(dw), b, n, dlogit.shape[-1],
                                        _stream()), "cd360_nerf_nviews_geo_bwd")
    return dw


@_op("loss")
def diffusion_loss(eps, x_noisy, target, sigma, mask, coef, *, ldd=64):
    """eps fp32 tokens [b*hw, 4]; x_noisy/target fp32 [b,4,h,w]; mask fp32 [b,1,h,w] | None.
    Returns (loss [b], mask_sum [b], deps bf16 [b*hw, ldd])."""
    lib = _lib.load()
    _req(eps, f32, "eps")
    _req(x_noisy, f32, "x_noisy")
    _req(target, f32, "target")
    b = x_noisy.shape[0]
    hw = x_noisy.shape[2] * x_noisy.shape[3]
    dev = eps.device
    loss = torch.empty(b, device=dev, dtype=f32)
    msum = torch.empty(b, device=dev, dtype=f32)
    deps = torch.empty((b * hw, ldd), device=dev, dtype=bf16)
    check(lib.cd360_diffusion_loss(_ptr(eps), _ptr(x_noisy), _ptr(target), _ptr(sigma), _ptr(mask),
                                   float(coef), _ptr(loss), _ptr(msum), _ptr(deps), b, hw, ldd, _stream()),
          "cd360_diffusion_loss")
    return loss, msum, deps


@_op("loss")
def nerf_aux_loss(fg, alphas, rgb, op, mask_s, tgt, mask_sum, wfg, wbg, wrgb):
    """Returns (loss3 [b,3], dfg, dalphas, drgb | None)."""
    lib = _lib.load()
    b, hw = fg.shape[0], fg.shape[1]
    d = alphas.shape[2]
    dev = fg.device
    loss3 = torch.empty((b, 3), device=dev, dtype=f32)
    dfg = torch.empty((b, hw), device=dev, dtype=f32)
    dal = torch.empty((b, hw, d), device=dev, dtype=f32)
    drgb = None if rgb is None else torch.empty((b, hw, 3), device=dev, dtype=f32)
    check(lib.cd360_nerf_aux_loss(_ptr(fg), _ptr(alphas), _ptr(rgb), _ptr(op), _ptr(mask_s), _ptr(tgt),
                                  _ptr(mask_sum), _ptr(wfg), _ptr(wbg), _ptr(wrgb), _ptr(loss3), _ptr(dfg),
                                  _ptr(dal), _ptr(drgb), b, hw, d, _stream()), "cd360_nerf_aux_loss")
    return loss3, dfg, dal, drgb


@_op("other")
def resize_bilinear_aa(x, oh, ow, *, scale=1.0, shift=0.0):
    """fp32 [..., ih, iw] -> [..., oh, ow], torch's bilinear antialias resize."""
    lib = _lib.load()
    _req(x, f32, "x")
    ih, iw = x.shape[-2:]
    planes = x.numel() // (ih * iw)
    out = torch.empty((*x.shape[:-2], oh, ow), device=x.device, dtype=f32)
    check(lib.cd360_resize_bilinear_aa(_ptr(x), _ptr(out), planes, ih, iw, oh, ow, float(scale),
                                       float(shift), _stream()), "cd360_resize_bilinear_aa")
    return out


@_op("adamw")
def adamw_step(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, step=1,
               grad_scale=1.0):
    lib = _lib.load()
    for t, nme in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _req(t, f32, nme)
    check(lib.cd360_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps,
                               weight_decay, step, grad_scale, _stream()), "cd360_adamw_step")
    return p
