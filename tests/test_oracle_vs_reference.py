"""Oracle vs the reference's own modules imported in place (skipped where /root/reference does
not exist, e.g. on the GPU box — the committed golden vectors cover that case)."""
import pytest
import torch

from oracle import ref_harness as H
from oracle import sgm_oracle as O

pytestmark = pytest.mark.skipif(not H.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ns():
    return H.import_reference()


def test_param_names_and_shapes_match_reference_module(ns):
    for cfg in (dict(O.TINY_CFG), dict(O.TINY_CFG, image_cross_blocks=[])):
        model = ns.openaimodel.UNetModel(**cfg)
        ref = {k: tuple(v.shape) for k, v in model.state_dict().items() if "raymarcher" not in k}
        assert ref == O.param_shapes(cfg)


def test_sdxl_topology_counts():
    """70 transformer blocks, 17 ResBlocks, 12 FeatureNeRF blocks (3 @640, 9 @1280)."""
    lay = O.unet_layout(O.SDXL_CFG)
    layers = [l for blk in lay["input_blocks"] + [lay["middle_block"]] + lay["output_blocks"] for l in blk]
    assert sum(l[3] for l in layers if l[0] == "st") == 70
    assert sum(1 for l in layers if l[0] == "res") == 17
    pb = O.pose_block_prefixes(O.SDXL_CFG)
    assert len(pb) == 12 and sum(1 for p in pb if p[1] == 640) == 3 and sum(1 for p in pb if p[1] == 1280) == 9
    n_params = sum(int(torch.tensor(s).prod()) for s in O.param_shapes(dict(O.SDXL_CFG, image_cross_blocks=[])).values())
    assert abs(n_params / 1e6 - 2567.46) < 0.01


def test_feature_nerf_substages(ns):
    """FeatureNeRFEncoding + Raymarcher of the reference vs the restatement, one block."""
    torch.manual_seed(0)
    c, n, res, d = 64, 5, 8, 6
    mod = ns.nerf.NerfSDModule(mode="feature-nerf", out_channels=c, far_plane=2.0, num_samples=d,
                               rgb_predict=True, average=False, num_freqs=16, stratified=True,
                               imp_sampling_percent=0.9, near_plane=0.0).eval()
    torch.nn.init.normal_(mod.model.decoder.weight, std=0.1)
    sd = {"m." + k: v for k, v in mod.model.state_dict().items()}
    cams = torch.stack([O.lookat_cameras(n, seed=3), O.lookat_cameras(n, seed=4, target_azimuth=2.0)])
    xref = torch.randn(2, n, res * res, c)
    pose = [H.cameras_from_packed(cams[0]), H.cameras_from_packed(cams[1])]
    with torch.no_grad():
        feats, sig, dists, attn, rgb, _, _ = mod(pose, xref, None)
        f2, rgb2, sig2, dists2, attn2 = O.feature_nerf_encoding(sd, "m.", cams, xref, d, 2.0)
    assert (feats - f2).abs().max() < 2e-5
    assert (sig - sig2).abs().max() < 2e-5 and (rgb - rgb2).abs().max() < 2e-5
    assert (attn - attn2).abs().max() < 2e-6
    assert torch.equal(dists, dists2)
    ren = ns.nerf.VolRender()
    with torch.no_grad():
        a = ren(feats, torch.exp(sig), dists, return_weights_uniform=True, rgb=torch.sigmoid(rgb))
        b = O.vol_render(feats, torch.exp(sig), dists, torch.sigmoid(rgb))
    for u, v in zip((a[0], a[1], a[2], a[4]), b):
        assert (u - v).abs().max() < 1e-6


def test_guider_and_scalings(ns):
    g = ns.guiders.ScheduledCFGImgTextRef(scale=7.5, scale_im=3.5)
    x = torch.randn(2, 4, 8, 8)
    s = torch.tensor([3.0, 3.0])
    c = {"crossattn": torch.randn(2, 77, 16), "vector": torch.randn(2, 12)}
    uc = {"crossattn": torch.zeros(2, 77, 16), "vector": torch.randn(2, 12)}
    a = g.prepare_inputs(x, s, c, uc)
    b = O.guider_prepare_inputs(x, s, c, uc, 3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert all(torch.equal(a[2][k], b[2][k]) for k in c)
    y = torch.randn(6, 4, 8, 8)
    assert torch.equal(g(y, s), O.guider_combine(y, 3, 7.5, 3.5))
    g2 = ns.guiders.VanillaCFGImgRef(scale=7.5)
    assert torch.equal(g2(y, s), O.guider_combine(y, 2, 7.5))
    a2 = g2.prepare_inputs(x, s, c, uc)
    b2 = O.guider_prepare_inputs(x, s, c, uc, 2)
    assert all(torch.equal(a2[2][k], b2[2][k]) for k in c)


def test_reference_stream_forward(ns):
    """UNetModel.forward WITH input_ref (the training-step call shape, forward only): the
    reference's own un-patched modules vs the restatement, including the tokens the validation hook
    would store as `references` (diffusion.py:28-40 keeps block outputs whose fg_mask is None)."""
    torch.manual_seed(0)
    cfg = dict(O.TINY_CFG)
    L, n, b = 16, 3, 1
    sd = O.synthetic_state_dict(cfg, seed=0, latent=L, num_references=n + 1)
    model = H.build_reference_unet(ns, cfg, sd, patch_for_sampling=False)
    captured = {}

    def hook(name):
        def fn(module, inp, out):
            if isinstance(out, tuple) and out[1] is None:
                captured.setdefault(name, []).append(out[0].detach())
        return fn

    handles = []
    for name, module in model.named_modules():
        parts = name.split(".")
        if len(parts) > 1 and parts[-2] == "transformer_blocks" and hasattr(module, "pose_emb_layers"):
            handles.append(module.register_forward_hook(hook(name)))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(b, 4, L, L, generator=g)
    xr = torch.randn(b, n, 4, L, L, generator=g)
    ctx = torch.randn(b, 77, cfg["context_dim"], generator=g)
    ctxr = torch.randn(b * n, 77, cfg["context_dim"], generator=g)
    y = torch.randn(b, cfg["adm_in_channels"], generator=g)
    yr = torch.randn(b * n, cfg["adm_in_channels"], generator=g)
    t = torch.tensor([500])
    sig = torch.tensor([120])
    cams = O.lookat_cameras(n, seed=3)[None]
    pose = [H.cameras_from_packed(cams[0])]
    with torch.no_grad():
        eps, fg, alphas, rgb = model(x, timesteps=t, context=torch.cat([ctx, ctxr]), y=torch.cat([y, yr]),
                                     input_ref=xr, sigmas_ref=sig, pose=pose, mask_ref=None)
        (eps2, aux2), cap2 = O.unet_forward_with_reference_stream(sd, cfg, x, t, ctx, y, cams, xr, sig,
                                                                  ctxr, yr)
    for h in handles:
        h.remove()
    assert len(fg) == len(aux2) == len(cap2) == len(captured) > 0
    assert (eps - eps2).abs().max() < 2e-4 * max(1.0, float(eps.abs().max()))
    for name, lst in captured.items():
        assert len(lst) == 1
        assert (lst[0] - cap2[name + "."]).abs().max() < 2e-4 * max(1.0, float(lst[0].abs().max()))
    for f, (f2, a2, r2) in zip(fg, aux2):
        assert (f - f2.reshape(f.shape)).abs().max() < 1e-4
