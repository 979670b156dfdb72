"""Shape-only stand-ins for custom_diffusion360_b200.ops, used by the CPU tests of the HOST logic
(module wiring, state-dict keys, topology, buffer shapes, sampler control flow).  They validate
operand shapes / dtypes exactly like the kernels' argument checks do and return zero tensors of
the right shape — no arithmetic is emulated (numerics are covered by the `-m gpu` tests)."""
import contextlib

import torch

bf16, f32 = torch.bfloat16, torch.float32


class Fake:
    def __init__(self):
        self.calls = []

    def _log(self, name, **kw):
        self.calls.append((name, kw))

    def gemm(self, a, w, *, bias=None, row_bias=None, rows_per_group=0, residual=None, out=None,
             out_fp32=False, act=0, geglu=False, a1=None, block_n=0, max_ctas=0, lda=None, lda1=None,
             k0=None, k1=None, M=None, ln_stats=None, ln_colsum=None, ln_eps=1e-5, stats_out=None):
        assert a.dtype == bf16 and w.dtype == bf16
        M = a.shape[0]
        K = a.shape[1] + (a1.shape[1] if a1 is not None else 0)
        assert w.shape[1] == K, (w.shape, K)
        if a1 is not None:
            assert a1.shape[0] == M and a.shape[1] % 64 == 0
        N = w.shape[0]
        n_out = N // 2 if geglu else N
        if bias is not None:
            assert bias.dtype == f32 and bias.shape == (N,)
        if residual is not None:
            assert residual.shape == (M, n_out) and residual.dtype == bf16
        if row_bias is not None:
            assert row_bias.shape[1] == N and rows_per_group > 0
        if ln_stats is not None:
            assert ln_stats.shape == (M, K // 64, 2) and ln_colsum.shape == (N,) and a1 is None
        if stats_out is not None:
            assert stats_out.shape == (M, n_out // 64, 2) and n_out % 64 == 0
        self._log("gemm", M=M, N=N, K=K, geglu=geglu, ln=ln_stats is not None, stats=stats_out is not None)
        if out is not None:
            assert out.shape == (M, n_out)
            return out
        return torch.zeros(M, n_out, dtype=f32 if out_fp32 else bf16)

    def gemm_tn(self, a_t, w_t, *, bias=None, out=None, out_fp32=True, k_splits=None):
        assert a_t.dtype == bf16 and w_t.dtype == bf16 and a_t.shape[0] == w_t.shape[0]
        assert a_t.stride(1) == 1 and w_t.stride(1) == 1
        M, N = a_t.shape[1], w_t.shape[1]
        assert M % 8 == 0 and N % 8 == 0 and a_t.stride(0) % 8 == 0 and w_t.stride(0) % 8 == 0
        self._log("gemm_tn", M=M, N=N, K=a_t.shape[0])
        if out is not None:
            assert out.shape == (M, N)
            return out
        return torch.zeros(M, N, dtype=f32 if out_fp32 else bf16)

    def conv3x3(self, x, w, B, H, W, *, bias=None, row_bias=None, residual=None, out=None,
                out_fp32=False, block_n=0, max_ctas=0):
        assert x.dtype == bf16 and x.shape[0] == B * H * W and w.shape[1] == 9 * x.shape[1]
        assert x.shape[1] % 64 == 0 and (H & (H - 1)) == 0 and (W & (W - 1)) == 0
        N = w.shape[0]
        if row_bias is not None:
            assert row_bias.shape == (B, N)
        if residual is not None:
            assert residual.shape == (B * H * W, N)
        self._log("conv3x3", M=B * H * W, N=N, C=x.shape[1])
        return torch.zeros(B * H * W, N, dtype=f32 if out_fp32 else bf16)

    def geglu_pack_block(self, n):
        return 128 if n % 256 == 0 else 64

    def attention(self, q, k, v, batch, heads, nq, nkv, *, out=None, ldq=None, ldk=None, ldv=None):
        assert q.shape[0] == batch * nq and k.shape[0] == batch * nkv and v.shape[0] == batch * nkv
        assert q.shape[1] == heads * 64 and k.shape[1] == heads * 64
        self._log("attention", batch=batch, heads=heads, nq=nq, nkv=nkv)
        return torch.zeros(batch * nq, heads * 64, dtype=bf16)

    def groupnorm(self, x0, gamma, beta, batch, hw, *, x1=None, eps=1e-5, silu=True, out=None, workspace=None):
        c = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        assert x0.shape[0] == batch * hw and gamma.shape == (c,) and c % 32 == 0
        self._log("groupnorm", c=c, eps=eps, silu=silu)
        return torch.zeros(batch * hw, c, dtype=bf16)

    def layernorm(self, x, gamma, beta, *, eps=1e-5, out=None):
        assert x.dtype == bf16 and gamma.shape == (x.shape[1],)
        self._log("layernorm", rows=x.shape[0])
        return torch.zeros_like(x)

    def small_linear(self, x, w, bias=None, *, add=None, act_in=0, act_out=0, out=None):
        assert x.dtype == f32 and w.dtype == bf16 and w.shape[1] == x.shape[1]
        return torch.zeros(x.shape[0], w.shape[0], dtype=f32)

    def timestep_embedding(self, t, dim, *, out=None):
        assert t.dtype == f32
        return torch.zeros(t.shape[0], dim, dtype=f32)

    def im2col3x3_nchw(self, x, kpad, *, scale=None, out=None, batch=None):
        b = x.shape[0] if batch is None else batch
        if scale is not None:
            assert scale.shape == (b,)
        return torch.zeros(b * x.shape[2] * x.shape[3], kpad, dtype=bf16)

    def im2col3x3_s2(self, x, batch, h, w, *, out=None):
        assert x.shape[0] == batch * h * w
        return torch.zeros(batch * (h // 2) * (w // 2), 9 * x.shape[1], dtype=bf16)

    def upsample_nearest2x(self, x, batch, h, w, *, out=None):
        assert x.shape[0] == batch * h * w
        return torch.zeros(4 * batch * h * w, x.shape[1], dtype=bf16)

    def cfg_euler_step_dev(self, x, eps, n_img, rows, hw, sig, scale, scale_im, *, denoised_out=None):
        assert eps.shape == (rows * n_img * hw, 4) and sig.shape == (3,)
        self._log("cfg_euler")
        return x

    def cast_bf16(self, x, *, out=None):
        return x.to(bf16)

    def cast_f32(self, x, *, out=None):
        return x.float()

    def nhwc_to_nchw_f32(self, x, batch, hw, c, *, out=None):
        return x.float().view(batch, hw, c).permute(0, 2, 1).contiguous()

    def nchw_to_nhwc_bf16(self, x, *, out=None):
        b, c = x.shape[:2]
        return x.reshape(b, c, -1).permute(0, 2, 1).reshape(-1, c).to(bf16)

    def nerf_points(self, cams, xy, depths, w_nv_geo, b_nv, b, n, res, d, kpe):
        assert cams.shape == (b, n + 1, 16) and xy.shape == (res * res, 2) and depths.shape == (res * res, d)
        assert w_nv_geo.shape == (198,)
        P = b * n * res * res * d
        return (torch.zeros(P, kpe, dtype=bf16), torch.zeros(P, 4, dtype=torch.int32),
                torch.zeros(P, 4), torch.zeros(b, n, res * res * d))

    def nerf_combine(self, g, hpre, gidx, gwgt, vlogit, b, n, hw, d, c):
        assert g.shape == (b * n * hw, c + 8) and hpre.shape == (b * n * hw * d, c)
        return torch.zeros(b * hw * d, c, dtype=bf16), torch.zeros(b, n, hw * d)

    def nerf_volrender(self, feats, raw, dists, b, hw, d, c):
        assert feats.shape == (b * hw * d, c) and raw.shape == (b * hw * d, 4) and dists.shape == (hw, d)
        self._log("volrender", b=b, hw=hw)
        return torch.zeros(b * hw, c, dtype=bf16), torch.zeros(b, hw), torch.zeros(b, hw, d), torch.zeros(b, hw, 3)

    # ---- training step -----------------------------------------------------------------------------
    def attention_bwd(self, q, k, v, o, dout, batch, heads, nq, nkv, *, dq, dk=None, dv=None):
        inner = heads * 64
        assert q.shape == (batch * nq, inner) and k.shape == (batch * nkv, inner) and v.shape == k.shape
        assert o.shape == q.shape and dout.shape == q.shape and dq.shape == q.shape
        assert (dk is None) == (dv is None) and (dk is None or (dk.shape == k.shape and dv.shape == k.shape))
        self._log("attention_bwd", nq=nq, nkv=nkv, kv_grad=dk is not None)
        return dq, dk, dv

    def layernorm_bwd(self, x, gamma, dy, *, add=None, eps=1e-5, out=None):
        assert x.dtype == bf16 and dy.shape == x.shape and gamma.shape == (x.shape[1],)
        assert add is None or add.shape == x.shape
        self._log("layernorm_bwd")
        return torch.zeros_like(x)

    def groupnorm_bwd(self, x0, gamma, beta, dy, batch, hw, *, x1=None, add0=None, add1=None, eps=1e-5, silu=True):
        c0 = x0.shape[1]
        c1 = 0 if x1 is None else x1.shape[1]
        assert x0.shape[0] == batch * hw and dy.shape == (batch * hw, c0 + c1) and gamma.shape == (c0 + c1,)
        assert add0 is None or add0.shape == x0.shape
        assert add1 is None or (x1 is not None and add1.shape == x1.shape)
        self._log("groupnorm_bwd", c=c0 + c1)
        return torch.zeros_like(x0), (None if x1 is None else torch.zeros_like(x1))

    def geglu_fwd(self, raw, block=None):
        assert raw.dtype == bf16 and raw.shape[1] % 2 == 0
        return torch.zeros(raw.shape[0], raw.shape[1] // 2, dtype=bf16)

    def geglu_bwd(self, raw, dh, block=None):
        assert raw.shape == (dh.shape[0], 2 * dh.shape[1]) and raw.dtype == bf16
        return torch.zeros_like(raw)

    def add_bf16(self, a, b, *, out=None):
        assert a.shape == b.shape and a.dtype == bf16
        self._log("add")
        return torch.zeros_like(a)

    def silu_bwd(self, pre, dy, *, out=None):
        assert pre.dtype == f32 and dy.dtype == f32 and pre.shape == dy.shape
        self._log("silu_bwd")
        return torch.zeros_like(pre)

    def transpose_to_bf16(self, x, *, ld_out=None):
        rows, cols = x.shape
        return torch.zeros(cols, (rows + 7) // 8 * 8 if ld_out is None else ld_out, dtype=bf16)

    def pointwise_conv_nchw(self, x, w, bias=None, *, scale=1.0, out=None):
        assert x.dtype == f32 and w.dtype == f32 and w.shape[1] == x.shape[1] and max(w.shape) <= 8
        self._log("pointwise_conv", cin=x.shape[1], cout=w.shape[0], scale=scale)
        return torch.zeros(x.shape[0], w.shape[0], *x.shape[2:], dtype=f32)

    def softmax_rows(self, s, *, scale=1.0, out=None):
        assert s.dtype == f32 and s.dim() == 2 and s.shape[1] % 4 == 0
        assert out is None or (out.shape == s.shape and out.dtype == bf16)
        self._log("softmax_rows", rows=s.shape[0], n=s.shape[1], scale=scale)
        return out if out is not None else torch.zeros(s.shape, dtype=bf16)

    def colsum(self, x, *, out=None):
        assert out is None or out.shape == (x.shape[1],)
        return out if out is not None else torch.zeros(x.shape[1])

    def col2im3x3_s2(self, dcol, batch, h, w, c):
        assert dcol.shape == (batch * (h // 2) * (w // 2), 9 * c)
        return torch.zeros(batch * h * w, c, dtype=bf16)

    def upsample_nearest2x_bwd(self, g, batch, h, w):
        assert g.shape[0] == 4 * batch * h * w
        return torch.zeros(batch * h * w, g.shape[1], dtype=bf16)

    def nerf_volrender_bwd(self, feats, raw, dists, d_rendered, dfg, dalphas, drgb, b, hw, d, c):
        assert feats.shape == (b * hw * d, c) and d_rendered.shape == (b * hw, c) and raw.shape == (b * hw * d, 4)
        assert dfg is None or dfg.shape == (b, hw)
        assert dalphas is None or dalphas.shape == (b, hw, d)
        assert drgb is None or drgb.shape == (b, hw, 3)
        self._log("volrender_bwd", seeded=dfg is not None)
        return torch.zeros_like(feats), torch.zeros(b * hw * d, 8, dtype=bf16)

    def nerf_combine_bwd(self, g, hpre, gidx, gwgt, vlogit, ds, b, n, hw, d, c):
        assert ds.shape == (b * hw * d, c) and hpre.shape == (b * n * hw * d, c)
        return torch.zeros_like(hpre), torch.zeros(b, n, hw * d), torch.zeros(b * n * hw, c + 8)

    def nerf_nviews_geo_bwd(self, cams, dlogit, b, n):
        assert dlogit.shape[:2] == (b, n)
        return torch.zeros(198)

    def diffusion_loss(self, eps, x_noisy, target, sigma, mask, coef, *, ldd=64):
        b = x_noisy.shape[0]
        hw = x_noisy.shape[2] * x_noisy.shape[3]
        assert eps.shape == (b * hw, 4) and target.shape == x_noisy.shape and sigma.shape == (b,)
        assert mask is None or mask.numel() == b * hw
        self._log("diffusion_loss", coef=coef)
        return torch.ones(b), torch.ones(b), torch.zeros(b * hw, ldd, dtype=bf16)

    def nerf_aux_loss(self, fg, alphas, rgb, op, mask_s, tgt, mask_sum, wfg, wbg, wrgb):
        b, hw = fg.shape
        assert alphas.shape[:2] == (b, hw) and op.shape == (b, hw)
        assert rgb is None or (rgb.shape == (b, hw, 3) and mask_s.shape == (b, hw) and tgt.shape == (b, 3, hw))
        self._log("nerf_aux_loss")
        return (torch.ones(b, 3), torch.zeros(b, hw), torch.zeros_like(alphas),
                None if rgb is None else torch.zeros(b, hw, 3))

    def resize_bilinear_aa(self, x, oh, ow, *, scale=1.0, shift=0.0):
        return torch.zeros(*x.shape[:-2], oh, ow)

    def adamw_step(self, p, g, m, v, **kw):
        assert p.shape == g.shape == m.shape == v.shape
        self._log("adamw", step=kw.get("step"))
        return p


@contextlib.contextmanager
def patched_ops():
    """Swap every arithmetic entry of custom_diffusion360_b200.ops for its shape-only stand-in."""
    from custom_diffusion360_b200 import ops
    fake = Fake()
    saved = {}
    for name in dir(fake):
        if name.startswith("_") or name == "calls":
            continue
        saved[name] = getattr(ops, name)
        setattr(ops, name, getattr(fake, name))
    try:
        yield fake
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
