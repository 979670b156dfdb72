"""Oracle (oracle/sgm_oracle.py) against the committed golden vectors, which were produced by
the reference's own modules (tests/golden/make_golden.py, reference commit 1a23f97).
Runs anywhere (CPU, no reference checkout needed).  fp32 vs fp32: tolerance 2e-5 absolute on
O(1) tensors (different op order / fused vs unfused kernels inside torch), exact for the schedule."""
import os

import pytest
import torch

from oracle import sgm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_unet_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu")


def _inputs(cfg, gold):
    inp = O.synthetic_inputs(cfg, gold["latent"], n_img=1, seed=0, n_views=gold["n_views"])
    c = {"crossattn": inp["crossattn"], "vector": inp["vector"]}
    uc = {"crossattn": torch.zeros_like(inp["crossattn"]), "vector": inp["vector"].clone()}
    uc["vector"][:, : cfg["adm_in_channels"] // 2] = 0
    return inp, c, uc


def test_sigma_schedule_bit_exact(gold):
    assert torch.equal(O.legacy_ddpm_sigmas(50), gold["sigmas_50"])
    assert torch.equal(O.legacy_ddpm_sigmas(1000, do_append_zero=False, flip=True), gold["sigmas_1000_flip"])
    assert torch.equal(O.legacy_ddpm_sigmas(gold["steps"]), gold["sigmas_%d" % gold["steps"]])


def test_timestep_embedding(gold):
    assert torch.equal(O.timestep_embedding(gold["t_emb_in"], 320), gold["t_emb_320"])


def test_unet_pose_off(gold):
    cfg = dict(O.TINY_CFG, image_cross_blocks=[])
    sd = O.synthetic_state_dict(cfg, seed=1)
    inp, c, _ = _inputs(dict(O.TINY_CFG), gold)
    with torch.no_grad():
        eps, aux = O.unet_forward(sd, cfg, inp["x"], torch.tensor([500]), c["crossattn"], c["vector"])
    assert aux == []
    assert (eps - gold["unet_eps_pose_off"]).abs().max() < 2e-5


def test_unet_pose_on_and_cache(gold):
    cfg = dict(O.TINY_CFG)
    L, nv = gold["latent"], gold["n_views"]
    sd = O.synthetic_state_dict(cfg, seed=0, latent=L, num_references=nv + 1)
    inp, c, uc = _inputs(cfg, gold)
    x3 = torch.cat([inp["x"]] * 3)
    ctx3 = torch.cat([uc["crossattn"], uc["crossattn"], c["crossattn"]])
    y3 = torch.cat([uc["vector"], uc["vector"], c["vector"]])
    t3 = torch.tensor([500, 500, 500])
    cams = inp["cams"][0][None].expand(3, -1, -1)
    cache = {}
    with torch.no_grad():
        eps, aux = O.unet_forward(sd, cfg, x3, t3, ctx3, y3, cams=cams, choices=list(range(nv)), cache=cache)
        eps2, aux2 = O.unet_forward(sd, cfg, 0.9 * x3, t3 - 100, ctx3, y3, cams=cams,
                                    choices=list(range(nv)), cache=cache)
    assert (eps - gold["unet_eps_step0"]).abs().max() < 2e-5
    assert (eps2 - gold["unet_eps_cached"]).abs().max() < 2e-5
    assert aux2 == [] and len(aux) == len(gold["fg_masks"]) == len(O.pose_block_prefixes(cfg))
    for (fg, al, rgb), gfg, gal, grgb in zip(aux, gold["fg_masks"], gold["alphas"], gold["rgbs"]):
        assert (fg - gfg).abs().max() < 2e-6
        assert (al - gal).abs().max() < 2e-6
        assert (rgb - grgb).abs().max() < 2e-6


def test_guidance_rows_structure_in_the_reference_golden(gold):
    """What may and what may not be shared between the three guidance rows of sample.py (pins the premise of
    FusedGuidedStep's row classes on the REFERENCE's own outputs): the rows are (uc text + null references,
    uc text + real references, c text + real references) — sample.py:85-96, guiders.py:114-128.  Rows 1 and 2
    therefore see the same cameras and reference tokens: their FeatureNeRF opacity maps (computed before any
    text enters: alphas come from the decoder's sigma, reference_attn order :571-598 — the text attention runs
    on the features, the densities are decoded from the encoding alone) are identical; row 0 differs; and the
    eps rows are pairwise different (no two rows can be merged)."""
    eps = gold["unet_eps_step0"]
    assert eps.shape[0] == 3
    for i, j in ((0, 1), (0, 2), (1, 2)):
        assert (eps[i] - eps[j]).abs().max() > 1e-3, (i, j)
    for al in gold["alphas"]:
        assert (al[1] - al[2]).abs().max() < 1e-6     # same encoding (the reference's batched matmuls differ by an ulp)
        assert (al[0] - al[1]).abs().max() > 1e-2     # null vs real references


def test_guided_euler_sampler(gold):
    """EulerEDMSampler + DiscreteDenoiser + ScheduledCFGImgTextRef around the UNet, 4 steps."""
    cfg = dict(O.TINY_CFG)
    L, nv = gold["latent"], gold["n_views"]
    sd = O.synthetic_state_dict(cfg, seed=0, latent=L, num_references=nv + 1)
    inp, c, uc = _inputs(cfg, gold)
    cams = inp["cams"][0][None].expand(3, -1, -1)
    cache = {}
    den = O.DiscreteDenoiserOracle()

    def network(x, c_noise, cond):
        return O.unet_forward(sd, cfg, x, c_noise, cond["crossattn"], cond["vector"], cams=cams,
                              choices=list(range(nv)), cache=cache) + (None, None)

    def net4(x, c_noise, cond):
        eps, aux = O.unet_forward(sd, cfg, x, c_noise, cond["crossattn"], cond["vector"], cams=cams,
                                  choices=list(range(nv)), cache=cache)
        return eps, None, None, None

    denoise_fn = lambda x, s, cc: den(net4, x, s, cc)[0]
    with torch.no_grad():
        out = O.euler_edm_sample(denoise_fn, inp["x"].clone(), c, uc, gold["steps"], rows=3,
                                 scale=7.5, scale_im=3.5)
    ref = gold["sample_final"]
    assert (out - ref).abs().max() < 1e-4 * max(1.0, float(ref.abs().max()))


def test_training_oracle_vs_reference_golden():
    """oracle/train_oracle.py against tests/golden/train_step_golden.pt — loss terms and pose
    gradients produced by the reference's own training code (tests/golden/make_train_golden.py),
    replaying the random draws it made (stratified jitter on, b = 2)."""
    import os

    from oracle import train_oracle as T
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "train_step_golden.pt"), map_location="cpu")
    case = gold["case"]
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=case["weights_seed"])
    batch = T.synthetic_train_batch(cfg, case["latent"], n_views=case["n_views"], b=case["b"], seed=case["batch_seed"],
                                    image=case["image"])
    batch["drop_im"] = torch.tensor(case["drop_im"])
    batch["rand"] = gold["rand"]
    total, terms, grads = T.training_gradients(sd, cfg, batch)
    assert abs(float(total) - gold["total"]) <= 1e-4 * max(1.0, abs(gold["total"]))
    for k, v in gold["terms"].items():
        assert abs(float(terms[k]) - v) <= 1e-4 * max(1.0, abs(v)), k
    assert set(grads) == set(gold["grads"])
    for k, ref in gold["grads"].items():
        g = grads[k]
        if k.endswith("nviews.bias"):      # sum over views of a softmax gradient: analytically zero, roundoff only
            assert float(g.abs().max()) <= 1e-7 and ref["norm"] <= 1e-7
            continue
        flat = g.reshape(-1)
        step = max(1, flat.numel() // 256)
        smp = flat[::step][:256]
        scale = max(float(ref["sample"].abs().max()), 1e-3 * ref["norm"], 1e-9)
        assert float((smp - ref["sample"]).abs().max()) <= 2e-3 * scale, k
        assert abs(float(g.norm()) - ref["norm"]) <= 2e-3 * max(ref["norm"], 1e-9), k


def test_vae_decoder_oracle_vs_reference_golden():
    """oracle/vae_oracle.py against the output of the reference's own Decoder (+ post_quant_conv and
    1/scale_factor) committed by tests/golden/make_vae_golden.py; the fixture is stored in fp16."""
    import os

    from oracle import vae_oracle as V
    from tests.golden.make_vae_golden import latent
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vae_decoder_golden.pt"), weights_only=False)
    cfg = dict(V.TINY_VAE_CFG)
    sd = V.synthetic_state_dict(cfg, seed=g["seed_w"])
    z = latent(g["seed_z"], g["batch"], g["latent"])
    img = V.decode_first_stage(sd, cfg, z, g["scale_factor"])
    ref = g["image"].float()
    assert img.shape == ref.shape == (g["batch"], 3, 8 * g["latent"], 8 * g["latent"])
    assert float((img - ref).abs().max()) <= 2e-3 * float(ref.abs().max())       # fp16 storage of the fixture
