"""Per-kernel numerics of the training-step kernels (csrc/train.cu, csrc/attention_bwd.cu): each
backward / loss / optimiser kernel, called through the C ABI, against torch.autograd of a plain fp32
PyTorch statement of the same op on the same (bf16-rounded) inputs.

Tolerance: gradients leave the kernels as bf16 (2^-9 relative rounding) after fp32 arithmetic on
bf16-rounded operands, so per tensor we require rel_rms <= 1e-2 and max |err| <= 4 % of max |ref|
(stated per test where an op needs more: attention recomputes P from bf16 scores)."""
import math

import pytest
import torch
import torch.nn.functional as F

gpu = pytest.mark.gpu
bf16 = torch.bfloat16


def _dev():
    return torch.device("cuda:0")


def _rt(x):
    xb = x.to(bf16)
    return xb, xb.float()


def _check(out, ref, rel_rms=1e-2, max_frac=4e-2, what=""):
    out, ref = out.float(), ref.float()
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    assert torch.isfinite(out).all(), what
    denom = ref.norm().clamp_min(1e-12)
    rr = float((out - ref).norm() / denom)
    mx = float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
    assert rr <= rel_rms and mx <= max_frac, f"{what}: rel_rms {rr:.4g} (tol {rel_rms}), max err / max|ref| {mx:.4g} (tol {max_frac})"


@gpu
@pytest.mark.parametrize("rows,c", [(64, 64), (1000, 640), (777, 1280), (33, 128)])
def test_layernorm_bwd(rows, c):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(0)
    x, xf = _rt(torch.randn(rows, c, device=_dev()) * 2 + 0.3)
    dy, dyf = _rt(torch.randn(rows, c, device=_dev()))
    add, addf = _rt(torch.randn(rows, c, device=_dev()))
    gamma = 1 + 0.1 * torch.randn(c, device=_dev())
    beta = 0.1 * torch.randn(c, device=_dev())
    xr = xf.clone().requires_grad_(True)
    F.layer_norm(xr, (c,), gamma, beta, 1e-5).backward(dyf)
    _check(ops.layernorm_bwd(x, gamma, dy), xr.grad, what="ln bwd")
    _check(ops.layernorm_bwd(x, gamma, dy, add=add), xr.grad + addf, what="ln bwd + add")


@gpu
@pytest.mark.parametrize("B,HW,c0,c1,silu,eps", [
    (2, 64, 64, 0, True, 1e-5),
    (1, 1024, 320, 0, True, 1e-5),      # cg = 10: 16-byte vectors straddle groups
    (3, 256, 640, 320, True, 1e-5),     # decoder concat
    (1, 256, 1280, 0, False, 1e-6),     # SpatialTransformer.norm
    (2, 100, 128, 64, False, 1e-5),
    (1, 4096, 320, 0, True, 1e-5),      # level 0 of the training step: 8-CTA clusters, 512 rows each
    (1, 900, 128, 64, True, 1e-5),      # rows not divisible by the cluster size (last rank: 109 of 113)
])
def test_groupnorm_bwd(B, HW, c0, c1, silu, eps):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(1)
    dev = _dev()
    c = c0 + c1
    x0, x0f = _rt(torch.randn(B * HW, c0, device=dev) * 1.5 + 0.2)
    x1 = x1f = None
    if c1:
        x1, x1f = _rt(torch.randn(B * HW, c1, device=dev))
    dy, dyf = _rt(torch.randn(B * HW, c, device=dev))
    addfull, addf = _rt(torch.randn(B * HW, c, device=dev))
    gamma = 1 + 0.1 * torch.randn(c, device=dev)
    beta = 0.1 * torch.randn(c, device=dev)
    xcat = (x0f if not c1 else torch.cat([x0f, x1f], 1)).reshape(B, HW, c).permute(0, 2, 1).clone().requires_grad_(True)
    y = F.group_norm(xcat, 32, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    y.backward(dyf.reshape(B, HW, c).permute(0, 2, 1))
    ref = xcat.grad.permute(0, 2, 1).reshape(B * HW, c)
    dx0, dx1 = ops.groupnorm_bwd(x0, gamma, beta, dy, B, HW, x1=x1, eps=eps, silu=silu)
    _check(dx0, ref[:, :c0], what="gn bwd x0")
    if c1:
        _check(dx1, ref[:, c0:], what="gn bwd x1")
    # accumulate a strided `add` (column slices of one buffer)
    dx0, dx1 = ops.groupnorm_bwd(x0, gamma, beta, dy, B, HW, x1=x1, eps=eps, silu=silu,
                                 add0=addfull[:, :c0], add1=addfull[:, c0:] if c1 else None)
    _check(dx0, ref[:, :c0] + addf[:, :c0], what="gn bwd x0 + add")
    if c1:
        _check(dx1, ref[:, c0:] + addf[:, c0:], what="gn bwd x1 + add")


@gpu
@pytest.mark.parametrize("rows,f,block", [(100, 256, None), (777, 2560, 128), (64, 512, 64)])
def test_geglu_bwd(rows, f, block):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(2)
    dev = _dev()
    raw, rawf = _rt(torch.randn(rows, 2 * f, device=dev) * 1.5)
    dh, dhf = _rt(torch.randn(rows, f, device=dev))
    blk = f if block is None else block
    # column order of the packed layout: tiles of [blk a-columns | blk gate-columns]
    t = rawf.reshape(rows, f // blk, 2, blk)
    a = t[:, :, 0].reshape(rows, f).clone().requires_grad_(True)
    g = t[:, :, 1].reshape(rows, f).clone().requires_grad_(True)
    (a * F.gelu(g)).backward(dhf)
    ref = torch.stack([a.grad.reshape(rows, f // blk, blk), g.grad.reshape(rows, f // blk, blk)], 2).reshape(rows, 2 * f)
    _check(ops.geglu_bwd(raw, dh, block), ref, what="geglu bwd")


@gpu
@pytest.mark.parametrize("batch,heads,nq,nkv,self_attn", [
    (1, 2, 64, 64, True),
    (2, 3, 200, 200, True),      # ragged tiles
    (1, 10, 1024, 1024, True),   # level-1 self-attention of the 64x64 training latents
    (2, 4, 300, 77, False),      # text cross-attention: dq only
    (1, 5, 1536, 77, False),     # FeatureNeRF samples x text
    (1, 5, 4096, 77, False),     # ... long enough for the query-split dK / dV kernel
])
def test_attention_bwd(batch, heads, nq, nkv, self_attn):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(3)
    dev = _dev()
    inner = heads * 64
    if self_attn:
        qkv, qkvf = _rt(torch.randn(batch * nq, 3 * inner, device=dev))
        q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
        qf, kf, vf = qkvf[:, :inner], qkvf[:, inner:2 * inner], qkvf[:, 2 * inner:]
    else:
        q, qf = _rt(torch.randn(batch * nq, inner, device=dev))
        kv, kvf = _rt(torch.randn(batch * nkv, 2 * inner, device=dev))
        k, v, kf, vf = kv[:, :inner], kv[:, inner:], kvf[:, :inner], kvf[:, inner:]
    do, dof = _rt(torch.randn(batch * nq, inner, device=dev))

    def split(t, n):
        return t.reshape(batch, n, heads, 64).permute(0, 2, 1, 3).clone().requires_grad_(True)

    qr, kr, vr = split(qf, nq), split(kf, nkv), split(vf, nkv)
    o_ref = F.scaled_dot_product_attention(qr, kr, vr)
    o_ref.backward(dof.reshape(batch, nq, heads, 64).permute(0, 2, 1, 3))
    merge = lambda t, n: t.permute(0, 2, 1, 3).reshape(batch * n, inner)
    o = ops.attention(q, k, v, batch, heads, nq, nkv)
    if self_attn:
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(q, k, v, o, do, batch, heads, nq, nkv, dq=dqkv[:, :inner],
                          dk=dqkv[:, inner:2 * inner], dv=dqkv[:, 2 * inner:])
        _check(dqkv[:, :inner], merge(qr.grad, nq), rel_rms=2e-2, max_frac=6e-2, what="dq")
        _check(dqkv[:, inner:2 * inner], merge(kr.grad, nkv), rel_rms=2e-2, max_frac=6e-2, what="dk")
        _check(dqkv[:, 2 * inner:], merge(vr.grad, nkv), rel_rms=2e-2, max_frac=6e-2, what="dv")
    else:
        dq = torch.empty_like(q)
        ops.attention_bwd(q, k, v, o, do, batch, heads, nq, nkv, dq=dq)
        _check(dq, merge(qr.grad, nq), rel_rms=2e-2, max_frac=6e-2, what="dq (cross)")
        # conditioning gradients: dK / dV of the text context (strided column slices of one K|V buffer);
        # nq = 1536 takes the plain kernel, 4096 queries x 77 keys the query-split one
        dkv = torch.zeros_like(kv)
        dq2 = torch.empty_like(q)
        ops.attention_bwd(q, k, v, o, do, batch, heads, nq, nkv, dq=dq2, dk=dkv[:, :inner], dv=dkv[:, inner:])
        assert torch.equal(dq2, dq)
        _check(dkv[:, :inner], merge(kr.grad, nkv), rel_rms=2e-2, max_frac=6e-2, what="dk (cross)")
        _check(dkv[:, inner:], merge(vr.grad, nkv), rel_rms=2e-2, max_frac=6e-2, what="dv (cross)")


@gpu
def test_transpose_colsum_add():
    from custom_diffusion360_b200 import ops
    torch.manual_seed(4)
    dev = _dev()
    x, xf = _rt(torch.randn(1003, 330, device=dev))
    t = ops.transpose_to_bf16(x)
    assert t.shape == (330, 1008)
    assert torch.equal(t[:, :1003].float(), xf.t()) and float(t[:, 1003:].abs().max()) == 0.0
    x32 = torch.randn(77, 648, device=dev)
    t32 = ops.transpose_to_bf16(x32[:, :640])           # strided fp32 source
    assert torch.equal(t32[:, :77], x32[:, :640].t().to(bf16))
    s = ops.colsum(x)
    _check(s, xf.sum(0), rel_rms=1e-5, max_frac=1e-4, what="colsum")
    y, yf = _rt(torch.randn(1003, 328, device=dev))
    z = ops.add_bf16(x[:, :328].contiguous(), y)
    assert torch.equal(z, (xf[:, :328] + yf).to(bf16))


@gpu
def test_weight_gradient_through_transposes():
    """dW = dY^T X as the forward GEMM over transposed operands (fp32 out, strided destination)."""
    from custom_diffusion360_b200 import ops
    torch.manual_seed(5)
    dev = _dev()
    M, N, K = 1536, 128, 328
    dy, dyf = _rt(torch.randn(M, N, device=dev))
    x, xf = _rt(torch.randn(M, K, device=dev))
    dw = torch.zeros(N, 2 * K, device=dev)
    ops.gemm(ops.transpose_to_bf16(dy), ops.transpose_to_bf16(x), out=dw[:, K:])
    _check(dw[:, K:], dyf.t() @ xf, rel_rms=1e-4, max_frac=1e-3, what="dW")
    assert float(dw[:, :K].abs().max()) == 0.0


@gpu
@pytest.mark.parametrize("B,H,W,C", [(1, 8, 8, 64), (2, 16, 32, 128)])
def test_resampling_bwd(B, H, W, C):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(6)
    dev = _dev()
    # stride-2 im2col backward = adjoint of the forward im2col kernel
    dcol, dcolf = _rt(torch.randn(B * (H // 2) * (W // 2), 9 * C, device=dev))
    x, xf = _rt(torch.randn(B * H * W, C, device=dev))
    col = ops.im2col3x3_s2(x, B, H, W).float()
    dx = ops.col2im3x3_s2(dcol, B, H, W, C)
    xr = xf.reshape(B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    unf = F.unfold(xr, 3, padding=1, stride=2)                         # [B, C*9, L] with (c, ky, kx) order
    unf = unf.reshape(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B * (H // 2) * (W // 2), 9 * C)
    assert torch.equal(unf.detach(), col)
    unf.backward(dcolf)
    _check(dx, xr.grad.permute(0, 2, 3, 1).reshape(B * H * W, C), what="col2im s2")
    # nearest x2 upsample backward
    g, gf = _rt(torch.randn(B * 4 * H * W, C, device=dev))
    ref = F.avg_pool2d(gf.reshape(B, 2 * H, 2 * W, C).permute(0, 3, 1, 2), 2) * 4
    _check(ops.upsample_nearest2x_bwd(g, B, H, W), ref.permute(0, 2, 3, 1).reshape(B * H * W, C), what="upsample bwd")


@gpu
@pytest.mark.parametrize("cin,cout,H", [(64, 128, 16), (128, 64, 8)])
def test_conv3x3_data_gradient(cin, cout, H):
    """dX of a stride-1 3x3 conv = the same implicit-GEMM kernel over the tap-flipped, channel-
    transposed weight pack (prepack.pack_conv3x3_bwd)."""
    from custom_diffusion360_b200 import ops
    from custom_diffusion360_b200.sgm.prepack import pack_conv3x3_bwd
    torch.manual_seed(7)
    dev = _dev()
    B = 2
    w = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin)).to(bf16).float()
    dy, dyf = _rt(torch.randn(B * H * H, cout, device=dev))
    x = torch.randn(B, cin, H, H, device=dev, requires_grad=True)
    F.conv2d(x, w, padding=1).backward(dyf.reshape(B, H, H, cout).permute(0, 3, 1, 2))
    dx = ops.conv3x3(dy, pack_conv3x3_bwd(w), B, H, H)
    _check(dx, x.grad.permute(0, 2, 3, 1).reshape(B * H * H, cin), what="conv dX")


@gpu
@pytest.mark.parametrize("b,hw,d,c", [(1, 64, 6, 128), (2, 256, 24, 640)])
def test_volrender_bwd(b, hw, d, c):
    from custom_diffusion360_b200 import ops
    torch.manual_seed(8)
    dev = _dev()
    feats, featsf = _rt(torch.randn(b * hw * d, c, device=dev))
    raw = torch.randn(b * hw * d, 4, device=dev) * 0.7
    dists = torch.full((hw, d), 2.0 / d, device=dev) * (1 + 0.2 * torch.rand(hw, d, device=dev))
    dren, drenf = _rt(torch.randn(b * hw, c, device=dev))
    dfg = torch.randn(b, hw, device=dev)
    dal = torch.randn(b, hw, d, device=dev)
    drgb = torch.randn(b, hw, 3, device=dev)

    class TruncExp(torch.autograd.Function):  # attention.py:192-208
        @staticmethod
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return torch.exp(x)

        @staticmethod
        def backward(ctx, g):
            return g * torch.exp(ctx.saved_tensors[0].clamp(max=15))

    fr = featsf.reshape(b, hw, d, c).clone().requires_grad_(True)
    rr = raw.reshape(b, hw, d, 4).clone().requires_grad_(True)
    sigma = TruncExp.apply(rr[..., 3:])
    dd = dists[None, :, :, None] * sigma
    alphas = 1 - torch.exp(-dd)
    trans = torch.cat([torch.zeros(b, hw, 1, 1, device=dev), torch.cumsum(dd[..., :-1, :], -2)], -2)
    wts = torch.nan_to_num(alphas * torch.exp(-trans))
    rendered = (wts * fr).sum(-2)
    fg = wts.sum(-2)
    rgb = (wts * torch.sigmoid(rr[..., :3])).sum(-2)
    loss = (rendered * drenf.reshape(b, hw, c)).sum() + (fg[..., 0] * dfg).sum() + (alphas[..., 0] * dal).sum() + (rgb * drgb).sum()
    loss.backward()
    dfeats, draw = ops.nerf_volrender_bwd(feats, raw, dists, dren, dfg, dal, drgb, b, hw, d, c)
    _check(dfeats, fr.grad.reshape(b * hw * d, c), what="volrender dfeats")
    _check(draw[:, :4], rr.grad.reshape(b * hw * d, 4), rel_rms=1.5e-2, what="volrender draw")
    assert float(draw[:, 4:].abs().max()) == 0.0
    # forward outputs of the same inputs (kernel pair consistency)
    ren2, fg2, al2, rgb2 = ops.nerf_volrender(feats, raw, dists, b, hw, d, c)
    _check(fg2, fg[..., 0].detach(), what="volrender fg")


@gpu
@pytest.mark.parametrize("b,n,res,d,c", [(1, 4, 8, 6, 128), (2, 3, 16, 24, 640)])
def test_nerf_combine_bwd(b, n, res, d, c):
    """Backward of the gather / SiLU / view-softmax combine against autograd of its plain statement."""
    from custom_diffusion360_b200 import ops
    torch.manual_seed(9)
    dev = _dev()
    hw = res * res
    ldg = c + 8
    g, gf = _rt(torch.randn(b * n * hw, ldg, device=dev))
    hpre, hpref = _rt(torch.randn(b * n * hw * d, c, device=dev))
    P = b * n * hw * d
    gidx = torch.randint(-1, hw, (P, 4), device=dev, dtype=torch.int32)
    gwgt = torch.rand(P, 4, device=dev) * (gidx >= 0)
    vlogit = torch.randn(b, n, hw * d, device=dev)
    ds, dsf = _rt(torch.randn(b * hw * d, c, device=dev))

    gr = gf.clone().requires_grad_(True)
    hr = hpref.clone().requires_grad_(True)
    vr = vlogit.clone().requires_grad_(True)
    G = gr.reshape(b, n, hw, ldg)
    idx = gidx.reshape(b, n, hw * d, 4).long().clamp_min(0)
    gathered = torch.zeros(b, n, hw * d, ldg, device=dev)
    for k in range(4):
        rows = torch.gather(G, 2, idx[..., k:k + 1].expand(-1, -1, -1, ldg))
        gathered = gathered + gwgt.reshape(b, n, hw * d, 4)[..., k:k + 1] * rows
    s = F.silu(hr.reshape(b, n, hw * d, c) + gathered[..., :c])
    a = torch.softmax(vr + gathered[..., c], dim=1)
    S = (a[..., None] * s).sum(1)
    (S * dsf.reshape(b, hw * d, c)).sum().backward()

    s_out, vsm = ops.nerf_combine(g, hpre, gidx, gwgt, vlogit, b, n, hw, d, c)
    _check(s_out, S.detach().reshape(b * hw * d, c), what="combine fwd")
    dhpre, dlogit, dg = ops.nerf_combine_bwd(g, hpre, gidx, gwgt, vlogit, ds, b, n, hw, d, c)
    _check(dhpre, hr.grad, what="dhpre")
    _check(dlogit, vr.grad, rel_rms=1e-3, what="dlogit")
    _check(dg[:, :c + 1], gr.grad[:, :c + 1], rel_rms=2e-3, what="dG")


@gpu
def test_nviews_geo_bwd():
    from custom_diffusion360_b200 import ops
    from oracle import sgm_oracle as O
    torch.manual_seed(10)
    dev = _dev()
    b, n, pts = 2, 4, 500
    cams = torch.stack([O.lookat_cameras(n, seed=i) for i in range(b)]).to(dev)
    dlogit = torch.randn(b, n, pts, device=dev)
    dlogit -= dlogit.mean(1, keepdim=True)   # softmax gradients sum to zero over the views
    w = torch.randn(198, device=dev, requires_grad=True)
    c = cams.cpu()
    R, T = c[..., :9].reshape(b, n + 1, 3, 3), c[..., 9:12]
    centers = O.camera_centers(c)
    o_ref_t = torch.einsum("bnj,bjk->bnk", centers[:, 1:], R[:, 0]) + T[:, 0, None, :]
    feat = torch.cat([o_ref_t, O.positional_encoding(o_ref_t, 16)], -1).to(dev)      # [b, n, 99]
    logit = (feat * w[99:]).sum(-1)[..., None].expand(b, n, pts)
    (logit * dlogit).sum().backward()
    dw = ops.nerf_nviews_geo_bwd(cams, dlogit, b, n)
    assert float(dw[:99].abs().max()) == 0.0
    _check(dw[99:], w.grad[99:], rel_rms=1e-3, max_frac=1e-2, what="nviews geo dW")


@gpu
def test_losses_and_resize():
    from custom_diffusion360_b200 import ops
    torch.manual_seed(11)
    dev = _dev()
    b, L = 2, 16
    hw = L * L
    eps = torch.randn(b * hw, 4, device=dev, requires_grad=True)
    x = torch.randn(b, 4, L, L, device=dev)
    noise = torch.randn(b, 4, L, L, device=dev)
    sigma = torch.tensor([0.7, 3.1], device=dev)
    xn = x + noise * sigma[:, None, None, None]
    mask = (torch.rand(b, 1, L, L, device=dev) > 0.3).float()
    for mk in (mask, None):
        eps.grad = None
        mo = xn - sigma[:, None, None, None] * eps.reshape(b, L, L, 4).permute(0, 3, 1, 2)
        le = sigma[:, None, None, None] ** -2 * (mo - x) ** 2
        ref = (le * mk).sum([1, 2, 3]) / (mk.sum([1, 2, 3]) + 1e-6) if mk is not None else le.reshape(b, -1).mean(1)
        (0.5 * ref.sum()).backward()
        loss, msum, deps = ops.diffusion_loss(eps.detach(), xn, x, sigma, mk, 0.5)
        _check(loss, ref.detach(), rel_rms=1e-5, max_frac=1e-4, what="l2 loss")
        _check(deps[:, :4], eps.grad, what="deps")
        assert float(deps[:, 4:].abs().max()) == 0.0
        if mk is not None:
            _check(msum, mk.sum([1, 2, 3]), rel_rms=1e-6, max_frac=1e-6, what="mask sum")
    # antialiased bilinear resize, down and up, against torch
    img = torch.rand(3, 2, 64, 48, device=dev)
    for oh, ow in ((16, 16), (8, 12), (96, 64), (64, 48)):
        ref = F.interpolate(img, size=(oh, ow), mode="bilinear", antialias=True)
        _check(ops.resize_bilinear_aa(img, oh, ow), ref, rel_rms=1e-5, max_frac=1e-4, what=f"resize {oh}x{ow}")
    _check(ops.resize_bilinear_aa(img, 16, 16, scale=0.5, shift=0.5),
           F.interpolate(img * 0.5 + 0.5, size=16, mode="bilinear", antialias=True), rel_rms=1e-5, max_frac=1e-4, what="resize affine")
    # FeatureNeRF supervision terms
    d = 6
    fg = (torch.rand(b, hw, device=dev) * 1.4 - 0.2).requires_grad_(True)
    al = torch.rand(b, hw, d, device=dev, requires_grad=True)
    rgb = torch.rand(b, hw, 3, device=dev, requires_grad=True)
    op = torch.rand(b, hw, device=dev) * (torch.rand(b, hw, device=dev) > 0.5)
    ms = torch.rand(b, hw, device=dev)
    tgt = torch.rand(b, 3, hw, device=dev)
    msum = torch.tensor([100.0, 57.0], device=dev)
    wfg, wbg, wrgb = (torch.rand(b, device=dev) for _ in range(3))
    lfg = ((fg.clamp(0, 1) - op) ** 2).mean(1)
    lbg = ((al - op[..., None]).abs() * (1 - op[..., None]) * (op[..., None] < 0.1)).mean([1, 2])
    lrgb = (((tgt - rgb.permute(0, 2, 1)) ** 2) * ms[:, None]).sum([1, 2]) / (msum + 1e-6)
    ((lfg * wfg).sum() + (lbg * wbg).sum() + (lrgb * wrgb).sum()).backward()
    loss3, dfg, dal, drgb = ops.nerf_aux_loss(fg.detach(), al.detach(), rgb.detach(), op, ms, tgt, msum, wfg, wbg, wrgb)
    _check(loss3, torch.stack([lfg, lbg, lrgb], 1).detach(), rel_rms=1e-5, max_frac=1e-4, what="aux losses")
    _check(dfg, fg.grad, rel_rms=1e-5, max_frac=1e-4, what="dfg")
    _check(dal, al.grad, rel_rms=1e-5, max_frac=1e-4, what="dalphas")
    _check(drgb, rgb.grad, rel_rms=1e-5, max_frac=1e-4, what="drgb")


@gpu
def test_adamw_matches_torch():
    from custom_diffusion360_b200 import ops
    torch.manual_seed(12)
    dev = _dev()
    p0 = torch.randn(10007, device=dev)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 4):
        g = torch.randn_like(p0)
        ref.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g, m, v, lr=1e-3, step=step)
        _check(p, ref.detach(), rel_rms=1e-6, max_frac=1e-5, what=f"adamw step {step}")
