"""Host-side logic on CPU: module tree / state-dict contract, topology, kernel-call sequence and
operand shapes of a full forward (with shape-only stand-ins for the kernels, tests/fake_ops.py),
the inference caching semantics, config factory, and that the real ops refuse CPU tensors."""
import pytest
import torch

from oracle import sgm_oracle as O
from tests.fake_ops import patched_ops

P = "custom_diffusion360_b200.sgm.modules.diffusionmodules."


def _model(cfg):
    from custom_diffusion360_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
    return UNetModel(**cfg).eval()


def test_state_dict_contract_tiny_and_pose_off():
    for cfg in (dict(O.TINY_CFG), dict(O.TINY_CFG, image_cross_blocks=[])):
        m = _model(cfg)
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if "raymarcher" not in k}
        assert mine == O.param_shapes(cfg)
        ray = [k for k in m.state_dict() if "raymarcher" in k]
        assert len(ray) == 5 * len(O.pose_block_prefixes(cfg))  # u, lengths, lengths_{center,upper,lower}


def test_class_and_attribute_names_sample_py_relies_on():
    m = _model(dict(O.TINY_CFG))
    sts = [x for x in m.modules() if x.__class__.__name__ == "SpatialTransformer"]
    bts = [x for x in m.modules() if x.__class__.__name__ == "BasicTransformerBlock"]
    assert len(sts) == 11 and len(bts) == 2 * 5 + 5 * 6
    for st in sts:
        for a in ("norm", "proj_in", "proj_out", "use_linear", "transformer_blocks", "image_cross", "poscontrol_interval"):
            assert hasattr(st, a)
    for bt in bts:
        for a in ("attn1", "attn2", "norm1", "norm2", "norm3", "ff", "disable_self_attn", "rendered_feat"):
            assert hasattr(bt, a)
    names = [n for n, _ in m.pose_blocks()]
    assert [n + "." for n in names] == [p for p, _, _ in O.pose_block_prefixes(O.TINY_CFG)]
    pose_params = [n for n, _ in m.named_parameters() if "pose" in n]
    assert len(pose_params) == 8 * len(names)  # pose_emb + plane_coefs(4) + nviews(2) + decoder


def test_forward_call_sequence_and_shapes_with_fake_kernels():
    cfg = dict(O.TINY_CFG)
    L, nv = 16, 4
    m = _model(cfg)
    refs = {p[:-1]: torch.randn(nv + 1, (L // ds) ** 2, c) for p, c, ds in O.pose_block_prefixes(cfg)}
    m.register_references(refs)
    m.set_reference_choices(list(range(nv)))
    x = torch.randn(3, 4, L, L)
    ctx, y = torch.randn(3, 77, cfg["context_dim"]), torch.randn(3, cfg["adm_in_channels"])
    cams = torch.randn(3, nv + 1, 16)
    with patched_ops() as fake, torch.no_grad():
        eps, fg, al, rgb = m(x, timesteps=torch.tensor([500, 500, 500]), context=ctx, y=y, pose=cams,
                             mask_ref=None, drop_im=None)
        n_first = len(fake.calls)
        eps2, fg2, _, _ = m(x, timesteps=torch.tensor([400, 400, 400]), context=ctx, y=y, pose=cams)
        n_second = len(fake.calls) - n_first
    assert eps.shape == (3, 4, L, L) and eps.dtype == torch.float32
    npose = len(O.pose_block_prefixes(cfg))
    assert len(fg) == len(al) == len(rgb) == npose and fg2 == []
    assert fg[0].shape == (3, (L // 2) ** 2, 1) and al[0].shape == (3, (L // 2) ** 2, cfg["num_samples"], 1)
    assert n_second < n_first  # cached rendered features: no FeatureNeRF kernels on later steps
    kinds = [c[0] for c in fake.calls[:n_first]]
    assert kinds.count("volrender") == npose
    assert kinds.count("conv3x3") == 2 * 17 + 2 + 1  # 17 ResBlocks x 2, 2 Upsample convs, out conv
    assert kinds.count("attention") == 2 * 40 + npose  # attn1+attn2 per block, + attn2 over samples
    m.clear_rendered_feat()
    assert all(b.rendered_feat is None for _, b in m.pose_blocks())


def test_sdxl_topology_counts_on_meta_device():
    with torch.device("meta"):
        m = _model(dict(O.SDXL_CFG))
    assert sum(p.numel() for n, p in m.named_parameters() if "pose" not in n) == 2567463684
    assert len(list(m.pose_blocks())) == 12
    assert len(m.resblocks()) == 17


def test_engine_factory_and_generic_sampler_control_flow():
    from custom_diffusion360_b200.sgm.models.diffusion import DiffusionEngine
    from custom_diffusion360_b200.sgm.util import instantiate_from_config
    cfg = dict(O.TINY_CFG, image_cross_blocks=[])
    engine = instantiate_from_config({"target": "custom_diffusion360_b200.sgm.models.diffusion.DiffusionEngine", "params": dict(
        network_config={"target": P + "openaimodel.UNetModel", "params": cfg},
        denoiser_config={"target": P + "denoiser.DiscreteDenoiser", "params": {
            "num_idx": 1000, "weighting_config": {"target": P + "denoiser_weighting.EpsWeighting"},
            "scaling_config": {"target": P + "denoiser_scaling.EpsScaling"},
            "discretization_config": {"target": P + "discretizer.LegacyDDPMDiscretization"}}},
        sampler_config={"target": P + "sampling.EulerEDMSampler", "params": {
            "num_steps": 3, "device": "cpu",
            "discretization_config": {"target": P + "discretizer.LegacyDDPMDiscretization"},
            "guider_config": {"target": P + "guiders.VanillaCFGImgRef", "params": {"scale": 7.5}}}})})
    assert isinstance(engine, DiffusionEngine)
    assert torch.equal(engine.denoiser.sigmas, O.legacy_ddpm_sigmas(1000, do_append_zero=False, flip=True))
    assert not any(p.requires_grad for _, p in engine.model.diffusion_model.named_parameters())
    c = {"crossattn": torch.randn(1, 77, cfg["context_dim"]), "vector": torch.randn(1, cfg["adm_in_channels"])}
    with patched_ops() as fake:
        out = engine.sample(c, uc=c, batch_size=1, num_steps=3, noise=torch.randn(1, 4, 16, 16), fused=False)
    assert out.shape == (1, 4, 16, 16)
    assert sum(1 for k, _ in fake.calls if k == "conv3x3") == 3 * (2 * 17 + 2 + 1)  # 3 steps


def test_unsupported_options_raise():
    from custom_diffusion360_b200.sgm.modules import attention as A
    with pytest.raises(NotImplementedError):
        A.MemoryEfficientCrossAttention(128, heads=4, dim_head=32)
    with pytest.raises(NotImplementedError):
        A.MemoryEfficientCrossAttention(128, heads=2, dim_head=64, add_lora=True)
    with pytest.raises(NotImplementedError):
        _model(dict(O.TINY_CFG, use_linear_in_transformer=False))


def test_real_ops_refuse_cpu_tensors():
    from custom_diffusion360_b200._lib import Cd360Error
    m = _model(dict(O.TINY_CFG, image_cross_blocks=[]))
    with pytest.raises(Cd360Error):
        m(torch.randn(1, 4, 16, 16), timesteps=torch.tensor([1]), context=torch.randn(1, 77, 128), y=torch.randn(1, 96))


def test_training_step_control_flow_with_fake_kernels():
    """Host logic of the training step (loss -> taped forward -> explicit backward walk -> flat
    AdamW) with shape-checking stand-ins for the kernels: every pose block gets a FeatureNeRF
    backward, skip-connection gradients are joined; with the conditioning gradients on (default) every
    attn2 backward — 1 per transformer block + 1 per FeatureNeRF block — also yields dK / dV, every
    ResBlock contributes to dL/d(vector), and the result has the conditioner outputs' shapes; with
    them off the walk stops at the first pose block."""
    from collections import Counter

    from tests import test_train_step_gpu as G
    from oracle import train_oracle as T
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=2)
    batch = T.synthetic_train_batch(cfg, 16, n_views=3, b=2, seed=5, image=48, jitter=True)
    n_pose = len(O.pose_block_prefixes(cfg))
    with patched_ops() as fake:
        eng = G._engine(cfg, sd, torch.device("cpu"))
        eng.global_step = 1
        opt = eng.configure_optimizers()
        assert len(opt.buckets) == n_pose
        names = [n for n, p in eng.model.diffusion_model.named_parameters() if p.requires_grad]
        assert names and all("pose" in n for n in names)
        opt.zero_grad()
        loss = eng.training_step(G._to_engine_batch(batch, torch.device("cpu")))
        assert set(eng.last_loss_dict) == {"loss", "loss_fg", "loss_bg", "loss_rgb"}
        opt.step()
        calls = Counter(k for k, _ in fake.calls)
        kv_grads = sum(1 for k, kw in fake.calls if k == "attention_bwd" and kw["kv_grad"] and kw["nkv"] == 77)
        unet = eng.model.diffusion_model
        n_blocks = sum(len(m.transformer_blocks) for m in unet.modules() if type(m).__name__ == "SpatialTransformer")
        assert kv_grads == n_blocks + n_pose and calls["silu_bwd"] == 2
        cg = eng.last_cond_grads
        bt = G._to_engine_batch(batch, torch.device("cpu"))
        assert cg["crossattn"].shape == bt["cond"]["crossattn"].shape and cg["vector"].shape == bt["cond"]["vector"].shape
        # conditioning gradients off: no dK / dV of the text context, shorter walk
        n_gn = calls["groupnorm_bwd"]
        fake.calls.clear()
        unet.cond_grad = False
        eng.training_step(bt)
        assert eng.last_cond_grads is None
        assert not any(k == "attention_bwd" and kw["kv_grad"] and kw["nkv"] == 77 for k, kw in fake.calls)
        assert Counter(k for k, _ in fake.calls)["groupnorm_bwd"] < n_gn
    assert calls["volrender"] == calls["volrender_bwd"] == n_pose
    assert calls["nerf_aux_loss"] == 2 * n_pose and calls["diffusion_loss"] == 2 and calls["adamw"] == 1
    assert eng.global_step == 2
    # parameters / gradients are views of the flat buffers, each 16-byte aligned
    for p_, o in zip(opt.flat.params, opt.flat.offsets):
        assert o % 4 == 0 and p_.data_ptr() == opt.flat.data.data_ptr() + 4 * o
        assert p_.grad.data_ptr() == opt.flat.grad.data_ptr() + 4 * o


def test_delta_checkpoint_round_trip(tmp_path):
    """Delta-checkpoint key contract (main.py:611-625 save, sgm/util.py:225-237 load): pose weights and
    `references` buffers, under `model.diffusion_model.*`, nothing else; loading restores them."""
    from tests import test_train_step_gpu as G
    from custom_diffusion360_b200.sgm import util as U
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=2)
    eng = G._engine(cfg, sd, torch.device("cpu"))
    unet = eng.model.diffusion_model
    names = [n for n, _ in unet.pose_blocks()]
    refs = {}
    for n, m in unet.pose_blocks():
        c = m.pose_emb_layers.weight.shape[0]
        refs[n] = torch.randn(5, 16, c)
    unet.register_references(refs)
    path = tmp_path / "delta.ckpt"
    with pytest.raises(ValueError, match="embed"):      # the reference loader indexes embed[0] / embed[1]
        U.save_delta_checkpoint(eng, path)
    U.save_delta_checkpoint(eng, path, embed=[torch.zeros(1, 8), torch.zeros(1, 8)])
    ck = torch.load(path, weights_only=False)
    delta = ck["delta_state_dict"]
    keys = [k for k in delta if k != "embed"]
    assert all(k.startswith("model.diffusion_model.") for k in keys)
    assert all(("pose" in k and "raymarcher" not in k) or "references" in k for k in keys)
    trainable = {"model.diffusion_model." + n for n, p in unet.named_parameters() if p.requires_grad}
    assert trainable <= set(keys)
    assert {f"model.diffusion_model.{n}.references" for n in names} <= set(keys)
    eng2 = G._engine(cfg, O.synthetic_state_dict(cfg, seed=7), torch.device("cpu"))
    base = {"model.diffusion_model." + k: v for k, v in O.synthetic_state_dict(cfg, seed=7).items()}
    missing, unexpected = U.load_checkpoints(eng2, base, delta)
    assert not missing and not unexpected
    u2 = eng2.model.diffusion_model
    for (n, p), (_, q) in zip(unet.named_parameters(), u2.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p, q), n
    for (n, m), (_, m2) in zip(unet.pose_blocks(), u2.pose_blocks()):
        assert torch.equal(m.references, m2.references), n


def test_camera_bin_loads_without_pytorch3d(tmp_path):
    """camera.bin is a pickle of pytorch3d PerspectiveCameras lists (main.py:1025-1029); the loader
    must read it with pytorch3d absent and return packed [N,16] rows."""
    import sys
    import types

    from custom_diffusion360_b200.sgm import util as U
    assert "pytorch3d" not in sys.modules
    pkg = types.ModuleType("pytorch3d")
    ren = types.ModuleType("pytorch3d.renderer")
    cams = types.ModuleType("pytorch3d.renderer.cameras")

    class PerspectiveCameras(torch.nn.Module):   # pickles like the real one: plain tensor attributes on an nn.Module
        def __init__(self, R, T, focal_length, principal_point):
            super().__init__()
            self.R, self.T, self.focal_length, self.principal_point = R, T, focal_length, principal_point
            self.in_ndc = True

    PerspectiveCameras.__module__ = "pytorch3d.renderer.cameras"
    PerspectiveCameras.__qualname__ = "PerspectiveCameras"
    cams.PerspectiveCameras = PerspectiveCameras
    sys.modules.update({"pytorch3d": pkg, "pytorch3d.renderer": ren, "pytorch3d.renderer.cameras": cams})
    try:
        g = torch.Generator().manual_seed(0)
        mk = lambda: PerspectiveCameras(torch.randn(1, 3, 3, generator=g), torch.randn(1, 3, generator=g),
                                        torch.rand(1, 2, generator=g) + 1, torch.zeros(1, 2))
        val, train = [mk() for _ in range(3)], [mk() for _ in range(20)]
        torch.save([val, train], tmp_path / "camera.bin")
    finally:
        for k in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.cameras"):
            sys.modules.pop(k)
    cv, ct = U.load_camera_bin(tmp_path / "camera.bin")
    assert "pytorch3d" not in sys.modules
    assert cv.shape == (3, 16) and ct.shape == (20, 16)
    assert torch.equal(ct[4, :9], train[4].R.reshape(-1)) and torch.equal(ct[4, 9:12], train[4].T[0])
    assert torch.equal(cv[1, 12:14], val[1].focal_length[0])
    choices = U.reference_choices(20, 8)
    assert choices == [int(x) for x in torch.linspace(0, 20 - 20 / 8, 8)]       # sample.py:275-278
    pose = U.sample_pose(cv, ct, 2, choices)
    assert pose.shape == (9, 16) and torch.equal(pose[0], cv[2]) and torch.equal(pose[3], ct[choices[2]])


def test_vae_decoder_host_logic_with_fake_kernels():
    """First-stage decode: state-dict contract of the decode-only AutoencoderKL,
    kernel-call sequence of one decode, and DiffusionEngine.decode_first_stage folding 1/scale_factor
    into the post_quant_conv call."""
    from collections import Counter

    from oracle import vae_oracle as V
    from custom_diffusion360_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    cfg = dict(V.TINY_VAE_CFG)
    vae = AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=cfg, lossconfig={"target": "torch.nn.Identity"},
                                        monitor="val/rec_loss").eval()
    assert {k: tuple(v.shape) for k, v in vae.state_dict().items()} == V.param_shapes(cfg)
    sd = V.synthetic_state_dict(cfg, seed=1)
    missing, ignored = vae.load_decode_state_dict(dict(sd, **{"encoder.conv_in.bias": torch.zeros(3)}))
    assert not missing and ignored == ["encoder.conv_in.bias"]
    with pytest.raises(KeyError):
        vae.load_decode_state_dict({"decoder.bogus": torch.zeros(1)})
    with pytest.raises(NotImplementedError):
        vae.encode(torch.zeros(1, 3, 64, 64))
    with patched_ops() as fake:
        img = vae.decode(torch.randn(2, 4, 16, 16), scale=1 / 0.13025)
    assert img.shape == (2, 3, 128, 128)
    calls = Counter(k for k, _ in fake.calls)
    n_res = 2 + 3 * len(cfg["ch_mult"])
    assert calls["conv3x3"] == 2 * n_res + (len(cfg["ch_mult"]) - 1) + 1      # res convs + upsample convs + conv_out
    assert calls["softmax_rows"] == 2                                          # one hw x hw score matrix per image
    pw = [kw for k, kw in fake.calls if k == "pointwise_conv"]
    assert len(pw) == 1 and abs(pw[0]["scale"] - 1 / 0.13025) < 1e-9
    sm = [kw for k, kw in fake.calls if k == "softmax_rows"][0]
    assert sm["rows"] == sm["n"] == 256 and abs(sm["scale"] - (cfg["ch"] * cfg["ch_mult"][-1]) ** -0.5) < 1e-9


def test_splitk_plan_heuristic():
    """Split-K dispatch (ops._splitk_plan): never for the sampling step's shapes (M >= 3072 rows: the
    tile grid already covers the GPU), on for the training step's small-M / large-K GEMMs and weight
    gradients, off below K = 2560 and for epilogues the finish kernel cannot apply."""
    from custom_diffusion360_b200 import ops
    plan = lambda M, N, K, plain=True: ops._splitk_plan(M, N, K, None, None, None, plain)
    for M, N, K in ((3072, 1280, 1280), (3072, 1280, 5120), (3072, 10240, 1280), (12288, 640, 2560), (49152, 320, 2880)):
        assert plan(M, N, K) == 1, (M, N, K)
    assert plan(256, 1280, 5120) == 7           # FF2 of the level-2 main stream: 20 single CTAs -> 140
    assert plan(256, 1280, 10240) == 7          # dX of FF1
    assert plan(1288, 1280, 24576) == 1         # 110 single-CTA tiles: one wave already (gemm_tcgen05.cu pick_config)
    assert plan(648, 640, 24576) == 4           # weight gradient [c + 8, c] over 24 576 sample rows: 30 tiles -> 120
    assert plan(256, 1280, 1280) == 1           # fixed launch cost dominates below K = 2560
    assert plan(256, 1280, 5120, plain=False) == 1
    assert plan(256, 1282, 5120) == 1           # finish kernel works on 4-column vectors


def test_featurenerf_row_classes():
    """FusedGuidedStep._set_row_classes: UNet rows that share one FeatureNeRF encoding = same cameras (bitwise)
    and same reference tokens (first row group: the null reference; the others: the real ones — the grouping
    rule of `context_ref_tokens` / sample.py:85-96, `batch % 3` included)."""
    import types
    from custom_diffusion360_b200.sgm.modules.diffusionmodules.sampling import FusedGuidedStep
    mk = lambda B, n: types.SimpleNamespace(dedup_rows=True, B=B, n_img=n, dev="cpu", _class_key=None, _classes=None)
    cams = torch.randn(4, 5, 16)
    o = mk(3, 1)                                   # sample.py car0: rows (u, ic, c) of one image
    FusedGuidedStep._set_row_classes(o, cams[:1])
    assert o._class_key == ((0, 1), (0, 1, 1)) and o._classes[0].tolist() == [0, 1]
    o = mk(12, 4)                                  # sweep unit: four prompts, one target camera
    FusedGuidedStep._set_row_classes(o, cams[:1].repeat(4, 1, 1))
    assert o._class_key == ((0, 4), (0,) * 4 + (1,) * 8)
    o = mk(12, 4)                                  # four different cameras: rows 1 and 2 of each image still pair up
    FusedGuidedStep._set_row_classes(o, cams)
    assert o._class_key == (tuple(range(8)), (0, 1, 2, 3, 4, 5, 6, 7, 4, 5, 6, 7))
    o = mk(2, 1)                                   # two-row guider: nothing to share
    FusedGuidedStep._set_row_classes(o, cams[:1])
    assert o._class_key == ((0, 1), (0, 1)) and o._classes is None
    o = mk(6, 3)                                   # two-row guider, three images: the reference's batch % 3 rule
    FusedGuidedStep._set_row_classes(o, cams[:1].repeat(3, 1, 1))      # makes rows 0-1 the "null" group
    assert o._class_key == ((0, 2), (0, 0, 1, 1, 1, 1))
    o = mk(3, 1)
    o.dedup_rows = False
    FusedGuidedStep._set_row_classes(o, cams[:1])
    assert o._classes is None


def test_engine_rejects_unbuilt_training_options():
    """Options whose silent acceptance would corrupt training: any trainkeys other than
    'pose' would put frozen SDXL weights under AdamW weight decay with zero gradients."""
    from tests import test_train_step_gpu as G
    from custom_diffusion360_b200.sgm.models.diffusion import DiffusionEngine
    cfg = dict(O.TINY_CFG)
    base = G._engine_config(cfg) if hasattr(G, "_engine_config") else None
    if base is None:
        pytest.skip("engine config helper not available")
    for bad in ({"trainkeys": "all"}, {"trainkeys": "poseattn"}, {"ckpt_path": "x.ckpt"},
                {"scheduler_config": {"target": "x"}}, {"use_ema": True}):
        with pytest.raises(NotImplementedError):
            DiffusionEngine(**{**base, **bad})


def test_product_general_conditioner_glue_vs_reference_golden():
    """GeneralConditioner of the PRODUCT (sgm/modules/encoders/modules.py) — the input_keys pairing,
    chunk(2) into main / reference halves, feature concatenation, batch-axis append and both force_*
    switches — against outputs of the reference's own GeneralConditioner on the same toy embedders
    (tests/golden/conditioner_golden.pt, made by tests/golden/make_conditioner_golden.py).  The toy text
    embedders are plain torch; ConcatTimestepEmbedderND's kernel is stood in for by the oracle formula."""
    import os
    import sys
    import types
    from oracle import conditioner_oracle as C
    from tests import test_oracle_conditioner as T
    from custom_diffusion360_b200.sgm.modules.encoders import modules as M

    toy = types.ModuleType("cd360_toy_product_embedders")

    class ToyText(M.AbstractEmbModel):
        def __init__(self, dim, mul, pooled_dim=0):
            super().__init__()
            self.dim, self.mul, self.pooled_dim = dim, mul, pooled_dim

        def forward(self, x):
            z = T.toy_text(x, self.dim, self.mul)
            return (z, T.toy_text(x, self.pooled_dim, 0.25)[:, 0]) if self.pooled_dim else z

    class ToySize(M.ConcatTimestepEmbedderND):
        def forward(self, x):
            return C.concat_timestep_embedder_nd(x, self.outdim)

    toy.ToyText, toy.ToySize = ToyText, ToySize
    sys.modules["cd360_toy_product_embedders"] = toy
    size = {"target": "cd360_toy_product_embedders.ToySize", "params": {"outdim": 8}}
    cond = M.GeneralConditioner([
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": "cd360_toy_product_embedders.ToyText",
         "params": {"dim": 6, "mul": 1.0}},
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": "cd360_toy_product_embedders.ToyText",
         "params": {"dim": 10, "mul": 0.5, "pooled_dim": 7}},
        dict(size, is_trainable=False, input_keys="original_size_as_tuple,original_size_as_tuple_ref"),
        dict(size, is_trainable=False, input_keys="crop_coords_top_left,crop_coords_top_left_ref")])
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "conditioner_golden.pt"), map_location="cpu")
    batch = T.toy_batch()
    out = cond(dict(batch), force_ref_zero_embeddings=False)
    assert torch.equal(out["crossattn"], gold["toy_crossattn"]) and torch.equal(out["vector"], gold["toy_vector"])
    emb = T.oracle_embedders()
    keys = [e["input_keys"] for e in emb]
    for force_ref in (False, True):
        c, uc = cond.get_unconditional_conditioning(dict(batch), force_uc_zero_embeddings=keys,
                                                    force_ref_zero_embeddings=force_ref)
        oc, ouc = C.get_unconditional_conditioning(emb, batch, None, keys, force_ref)
        for k in ("crossattn", "vector"):
            assert torch.equal(c[k], oc[k]) and torch.equal(uc[k], ouc[k]), (k, force_ref)
    part = cond(dict(batch), [["txt", "txt_ref"]], True)
    assert float(part["crossattn"].abs().max()) == 0.0 and float(part["vector"][:, 7:].abs().max()) > 0.0
