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


def _import_reference_training(ns):
    """The reference's loss / denoiser / samplers, imported in place.  loss.py pulls LPIPS and the
    conditioner at module import (loss.py:6-7); neither is used by the 'l2' branch, so both are
    stubbed as empty shells (the installed package set has no open_clip / kornia)."""
    import importlib
    import sys
    import types

    def shell(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m

    class _Stub(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    for name in ("sgm.modules.autoencoding", "sgm.modules.autoencoding.lpips", "sgm.modules.autoencoding.lpips.loss"):
        shell(name)
    shell("sgm.modules.autoencoding.lpips.loss.lpips", LPIPS=_Stub)
    shell("sgm.modules.encoders")
    shell("sgm.modules.encoders.modules", GeneralConditioner=_Stub)
    if "fsspec" not in sys.modules:
        try:
            import fsspec  # noqa: F401
        except Exception:
            shell("fsspec")
    ns.loss = importlib.import_module("sgm.modules.diffusionmodules.loss")
    ns.denoiser = importlib.import_module("sgm.modules.diffusionmodules.denoiser")
    ns.wrappers = importlib.import_module("sgm.modules.diffusionmodules.wrappers")
    return ns


def _reference_training_step(ns, cfg, sd, batch, seed, train_mode):
    """Loss terms and pose gradients from the reference's OWN training code path
    (StandardDiffusionLossImgRef -> DiscreteDenoiser -> OpenAIWrapper -> UNetModel, torch autograd),
    plus the random draws it made (replayed from the same generator state, in its call order)."""
    _import_reference_training(ns)
    disc = {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"}
    loss_fn = ns.loss.StandardDiffusionLossImgRef(
        sigma_sampler_config={"target": "sgm.modules.diffusionmodules.sigma_sampling.CubicSampling",
                              "params": {"num_idx": 1000, "discretization_config": disc}},
        sigma_sampler_config_ref={"target": "sgm.modules.diffusionmodules.sigma_sampling.DiscreteSampling",
                                  "params": {"num_idx": 50, "discretization_config": disc}})
    denoiser = ns.denoiser.DiscreteDenoiser(
        weighting_config={"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
        num_idx=1000, discretization_config=disc)
    unet = H.build_reference_unet(ns, cfg, sd, patch_for_sampling=False)
    unet.train(train_mode)
    for name, p in unet.named_parameters():       # trainkeys == 'pose' (diffusion.py:139-144)
        p.requires_grad = "pose" in name
    net = ns.wrappers.OpenAIWrapper(unet)
    b, n = batch["x_ref"].shape[:2]
    pose = [H.cameras_from_packed(batch["cams"][i]) for i in range(b)]
    # the conditioner's outputs are autograd leaves here: their gradients are what the reference's
    # text encoders receive for the `<new1>` token rows (diffusion.py:343-356)
    ca = batch["crossattn"].clone().requires_grad_(True)
    vec = batch["vector"].clone().requires_grad_(True)
    conditioner = lambda bt: {"crossattn": ca, "vector": vec}
    torch.manual_seed(seed)
    loss, loss_fg, loss_bg, loss_rgb = loss_fn(net, denoiser, conditioner, batch["x"], batch["rgb"], batch["x_ref"],
                                               pose, batch["mask"], batch.get("mask_ref"), batch["opacity"], {})
    # DiffusionEngine.forward (diffusion.py:221-236) with the yaml's lambdas, global_step > 0
    drop = batch["drop_im"]
    total = loss.mean()
    lf = (loss_fg.mean(1) * drop).sum() / (drop.sum() + 1e-12)
    lb = (loss_bg.mean(1) * drop).sum() / (drop.sum() + 1e-12)
    lr = (loss_rgb.mean(1) * drop).sum() / (drop.sum() + 1e-12)
    total = total + 10.0 * lf + 10.0 * lb + 5.0 * lr
    total.backward()
    grads = {k: p.grad.clone() for k, p in unet.named_parameters() if p.requires_grad}
    grads["cond.crossattn"], grads["cond.vector"] = ca.grad.clone(), vec.grad.clone()
    # replay the draws: CubicSampling torch.rand, randn_like(input), DiscreteSampling torch.randint,
    # randn_like(input_ref) (loss.py:147-170), randn_like(input_ref) (denoiser.py:31), then per pose
    # block in execution order: rand_like x2 (utils_cameraray.py:111-140), torch.rand (nerfsd:317-325)
    torch.manual_seed(seed)
    u = torch.rand((b,))
    rand = dict(sigma_idx=((1 - u ** 3) * 999).long(), noise=torch.randn_like(batch["x"]),
                sigma_ref_idx=torch.randint(0, 50, (b,)), noise_ref=torch.randn_like(batch["x_ref"]),
                noise_ref2=torch.randn_like(batch["x_ref"]))
    if train_mode and cfg.get("stratified"):
        L = batch["x"].shape[-1]
        jit = []
        for _, c, ds in O.pose_block_prefixes(cfg):
            res = L // ds
            rx = torch.rand(res + 1)
            ry = torch.rand(res + 1)
            jit.append(dict(xy_rand=(rx, ry), t_rand=torch.rand(res * res, cfg["num_samples"] + 1)))
        rand["jitter"] = jit
    terms = dict(loss=loss.mean().detach(), loss_fg=lf.detach(), loss_bg=lb.detach(), loss_rgb=lr.detach())
    return total.detach(), terms, grads, rand


@pytest.mark.parametrize("train_mode,mask_ref", [(False, False), (True, False), (True, True)])
def test_training_step_vs_reference(ns, train_mode, mask_ref):
    """The training oracle (oracle/train_oracle.py) against the reference's own training code:
    loss terms and the gradient of every trainable ('pose') parameter.  train_mode=True runs the
    reference UNet in .train() — stratified ray / depth jitter on (yaml: stratified: True);
    mask_ref=True passes the padding masks of the reference views (data_co3d.py:485 ->
    loss.py:154 -> nerfsd_pytorch3d.py:61-70), as every reference training batch does."""
    from oracle import train_oracle as T
    cfg = dict(O.TINY_CFG)
    L, n = 16, 3
    sd = O.synthetic_state_dict(cfg, seed=2)
    batch = T.synthetic_train_batch(cfg, L, n_views=n, b=1, seed=5, image=48, mask_ref=mask_ref)
    total_ref, terms_ref, grads_ref, rand = _reference_training_step(ns, cfg, sd, batch, seed=11, train_mode=train_mode)
    batch = dict(batch, rand=rand)
    total, terms, grads = T.training_gradients(sd, cfg, batch, cond_grads=True)
    assert abs(float(total) - float(total_ref)) <= 1e-4 * max(1.0, abs(float(total_ref)))
    # conditioning gradients: target rows carry signal, reference-view rows are exactly zero
    b_ = batch["x"].shape[0]
    for k in ("cond.crossattn", "cond.vector"):
        assert float(grads_ref[k][:b_].abs().max()) > 0 and float(grads_ref[k][b_:].abs().max()) == 0.0, k
        assert float(grads[k][b_:].abs().max()) == 0.0, k
    for k in ("loss", "loss_fg", "loss_bg", "loss_rgb"):
        assert abs(float(terms[k]) - float(terms_ref[k])) <= 1e-4 * max(1.0, abs(float(terms_ref[k]))), k
    assert set(grads) == set(grads_ref) and len(grads) > 0
    for k, g in grads_ref.items():
        scale = max(float(g.abs().max()), 1e-6)
        assert float((grads[k] - g).abs().max()) <= 2e-3 * scale, (k, float((grads[k] - g).abs().max()), scale)
    if mask_ref:   # the masks must matter, or this case pins nothing
        with torch.no_grad():
            total_nomask, _ = T.training_loss(sd, cfg, {k: v for k, v in batch.items() if k != "mask_ref"})
        assert abs(float(total_nomask) - float(total)) > 1e-3 * abs(float(total))


def test_vae_decoder_oracle_vs_reference():
    """oracle/vae_oracle.py against the reference's own Decoder module imported in place
    (sgm/modules/diffusionmodules/model.py:604-757): parameter names / shapes and the output."""
    import importlib

    import torch.nn.functional as F

    from oracle import vae_oracle as V
    H.install()
    m = importlib.import_module("sgm.modules.diffusionmodules.model")
    cfg = dict(V.TINY_VAE_CFG)
    dec = m.Decoder(**cfg).eval()
    assert type(dec.mid.attn_1).__name__ == "MemoryEfficientAttnBlock"
    ours = {k[len("decoder."):]: s for k, s in V.param_shapes(cfg).items() if k.startswith("decoder.")}
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == ours
    sd = V.synthetic_state_dict(cfg, seed=4)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = dec(F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
    out = V.autoencoder_decode(sd, cfg, z)
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_product_edm_sampler_with_churn_vs_reference(ns):
    """The PRODUCT EulerEDMSampler (generic path: sampler_step / __call__, host logic only) against the
    reference's own class with stochastic churn switched on (s_churn > 0: sigma_hat, the extra noise draw,
    the step from sigma_hat; reference sampling.py:96-136) — same torch generator state, a linear toy
    denoiser, bit-identical trajectory."""
    import importlib

    H.install()
    ref_sampling = importlib.import_module("sgm.modules.diffusionmodules.sampling")
    from custom_diffusion360_b200.sgm.modules.diffusionmodules import sampling as ours
    P = "custom_diffusion360_b200.sgm.modules.diffusionmodules."
    kw = dict(s_churn=20.0, s_tmin=0.05, s_tmax=10.0, s_noise=1.003, num_steps=12, device="cpu")
    ref = ref_sampling.EulerEDMSampler(
        discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"},
        guider_config={"target": "sgm.modules.diffusionmodules.guiders.VanillaCFGImgRef", "params": {"scale": 3.0}}, **kw)
    mine = ours.EulerEDMSampler(
        discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"},
        guider_config={"target": P + "guiders.VanillaCFGImgRef", "params": {"scale": 3.0}}, **kw)

    def denoiser(x, sigma, c):   # the reference's 4-tuple contract (denoiser.py:22-44)
        d = x * (1.0 / (1.0 + sigma.reshape(-1, 1, 1, 1) ** 2)) + 0.01 * c["vector"].reshape(-1, 1, 1, 1)
        return d, None, None, None

    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    c = {"crossattn": torch.randn(2, 77, 16, generator=g), "vector": torch.randn(2, 1, generator=g)}
    uc = {"crossattn": torch.zeros(2, 77, 16), "vector": torch.randn(2, 1, generator=g)}
    torch.manual_seed(11)
    a, _ = ref(denoiser, x0.clone(), c, uc=uc)
    torch.manual_seed(11)
    b, _ = mine(denoiser, x0.clone(), c, uc=uc)
    assert torch.equal(a, b)
    torch.manual_seed(11)
    c0, _ = ours.EulerEDMSampler(
        discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"},
        guider_config={"target": P + "guiders.VanillaCFGImgRef", "params": {"scale": 3.0}},
        **dict(kw, s_churn=0.0))(denoiser, x0.clone(), c, uc=uc)
    assert not torch.equal(b, c0)      # the churn must matter, or this case pins nothing
