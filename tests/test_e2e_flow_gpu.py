"""The whole `sample.py` flow through ONE DiffusionEngine of this package, as a user of the reference
would drive it (sample.py:140-200): token ids -> GeneralConditioner (both text towers + size embedders)
-> `get_unconditional_conditioning` -> `engine.sample` (fused guided Euler loop, FeatureNeRF pose
conditioning, CUDA graphs) -> `engine.decode_first_stage`, against the fp32 CPU oracle composed the same
way (conditioner_oracle -> sgm_oracle sampler / denoiser / guider / UNet -> vae_oracle).  Every stage has
its own parity test; this one pins the glue between them (row slicing of c / uc, dtype / device hand-offs,
scale_factor, state shared between calls)."""
import json
import os

import pytest
import torch

from oracle import conditioner_oracle as C
from oracle import sgm_oracle as O
from oracle import vae_oracle as V

gpu = pytest.mark.gpu
P = "custom_diffusion360_b200.sgm.modules.diffusionmodules."
E = "custom_diffusion360_b200.sgm.modules.encoders.modules."
METRICS = {}


def _record(name, ours, ref):
    ours, ref = ours.detach().float().cpu(), ref.detach().float().cpu()
    rel = float((ours - ref).norm() / ref.norm().clamp_min(1e-12))
    METRICS[name] = dict(rel_rms=rel, max_abs=float((ours - ref).abs().max()), ref_max=float(ref.abs().max()))
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "e2e_parity_metrics.json"), "w") as f:
        json.dump(METRICS, f, indent=1)
    return rel


@gpu
def test_sample_py_flow_tokens_to_image_vs_oracle():
    from tests.test_conditioner_gpu import TINY_CLIP, TINY_OC, _load, _tokens
    from custom_diffusion360_b200.sgm.models.diffusion import DiffusionEngine
    dev = torch.device("cuda:0")
    clip_cfg, oc_cfg = dict(TINY_CLIP), dict(TINY_OC)
    # UNet sized for what the tiny conditioner emits: crossattn 128 + 128, vector 48 + 6 x 8
    cfg = dict(O.TINY_CFG, context_dim=256, adm_in_channels=96)
    L, nv, steps, N = 16, 4, 3, 1
    sd = O.synthetic_state_dict(cfg, seed=0, latent=L, num_references=nv + 1)
    clip_sd = C.synthetic_state_dict(C.clip_param_shapes(clip_cfg), seed=5)
    oc_sd = C.synthetic_state_dict(C.open_clip_param_shapes(oc_cfg), seed=3)
    vae_cfg = dict(V.TINY_VAE_CFG)
    vae_sd = V.synthetic_state_dict(vae_cfg, seed=7)
    size_cfg = lambda k: {"is_trainable": False, "input_keys": f"{k},{k}_ref", "target": E + "ConcatTimestepEmbedderND",
                          "params": {"outdim": 8}}
    disc = {"target": P + "discretizer.LegacyDDPMDiscretization"}
    engine = DiffusionEngine(
        network_config={"target": P + "openaimodel.UNetModel", "params": cfg},
        denoiser_config={"target": P + "denoiser.DiscreteDenoiser", "params": {
            "num_idx": 1000, "weighting_config": {"target": P + "denoiser_weighting.EpsWeighting"},
            "scaling_config": {"target": P + "denoiser_scaling.EpsScaling"}, "discretization_config": disc}},
        sampler_config={"target": P + "sampling.EulerEDMSampler", "params": {
            "num_steps": steps, "discretization_config": disc,
            "guider_config": {"target": P + "guiders.ScheduledCFGImgTextRef", "params": {"scale": 7.5, "scale_im": 3.5}}}},
        conditioner_config={"target": E + "GeneralConditioner", "params": {"emb_models": [
            {"is_trainable": False, "input_keys": "txt,txt_ref", "target": E + "FrozenCLIPEmbedder",
             "params": {"layer": "hidden", "layer_idx": 11, "arch": clip_cfg}},
            {"is_trainable": False, "input_keys": "txt,txt_ref", "target": E + "FrozenOpenCLIPEmbedder",
             "params": {"arch": oc_cfg, "layer": "penultimate", "always_return_pooled": True, "legacy": False}},
            size_cfg("original_size_as_tuple"), size_cfg("crop_coords_top_left"), size_cfg("target_size_as_tuple")]}},
        first_stage_config={"target": "custom_diffusion360_b200.sgm.models.autoencoder.AutoencoderKLInferenceWrapper",
                            "params": {"embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": vae_cfg,
                                       "lossconfig": {"target": "torch.nn.Identity"}}},
        scale_factor=V.SDXL_SCALE_FACTOR)
    net = engine.model.diffusion_model
    net.load_state_dict({k: v for k, v in sd.items() if not k.endswith("references")}, strict=False)
    _load(engine.conditioner.embedders[0], clip_sd)
    _load(engine.conditioner.embedders[1], oc_sd)
    engine = engine.to(dev).eval()
    net.register_references({k: v.to(dev) for k, v in sd.items() if k.endswith("references")})
    engine.set_reference_choices(list(range(nv)))
    missing, _ = engine.init_first_stage().load_decode_state_dict(vae_sd)
    assert not missing

    # ---- the batch sample.py builds: one prompt for the target, the prompts of the nv reference views ----
    size = lambda v, n: torch.tensor([v]).repeat(n, 1)
    t1, t1r = _tokens(clip_cfg, N, 1), _tokens(clip_cfg, nv * N, 3)
    t2, t2r = _tokens(oc_cfg, N, 2), _tokens(oc_cfg, nv * N, 4)
    sizes = {"original_size_as_tuple": size([512.0, 512.0], N), "original_size_as_tuple_ref": size([512.0, 384.0], nv * N),
             "crop_coords_top_left": size([0.0, 16.0], N), "crop_coords_top_left_ref": size([8.0, 0.0], nv * N),
             "target_size_as_tuple": size([512.0, 512.0], N), "target_size_as_tuple_ref": size([512.0, 512.0], nv * N)}
    batch = dict(txt=(t1, t2), txt_ref=(t1r, t2r), **{k: v.to(dev) for k, v in sizes.items()})
    batch_o = dict(txt=(t1, t2), txt_ref=(t1r, t2r), **sizes)
    inp = O.synthetic_inputs(cfg, L, n_img=N, seed=0, n_views=nv)
    cams = inp["cams"][0]
    noise = inp["x"]

    # ---- this package (sample.py:150-196) ----
    keys = [e.input_keys for e in engine.conditioner.embedders]
    c, uc = engine.conditioner.get_unconditional_conditioning(batch, force_uc_zero_embeddings=keys,
                                                              force_ref_zero_embeddings=True)
    for k in c:                                             # sample.py:183-185
        c[k], uc[k] = c[k][:(nv + 1) * N].to(dev), uc[k][:(nv + 1) * N].to(dev)
    samples = engine.sample(c, shape=noise.shape[1:], uc=uc, batch_size=N, num_steps=steps, noise=noise.clone(),
                            pose=[cams] * 3, drop_im=None, mask_ref=None)
    engine.clear_rendered_feat()
    image = engine.decode_first_stage(samples)
    torch.cuda.synchronize()

    # ---- the oracle, composed the same way ----
    emb_o = C.sdxl_conditioner(clip_sd, clip_cfg, oc_sd, oc_cfg, size_dim=8)
    keys_o = [e["input_keys"] for e in emb_o]
    with torch.no_grad():
        c_o, uc_o = C.get_unconditional_conditioning(emb_o, batch_o, force_uc_zero_embeddings=keys_o,
                                                     force_ref_zero_embeddings=True)
        c_o = {k: v[:(nv + 1) * N] for k, v in c_o.items()}
        uc_o = {k: v[:(nv + 1) * N] for k, v in uc_o.items()}
        den = O.DiscreteDenoiserOracle()
        cache = {}
        cams3 = cams[None].expand(3, -1, -1)

        def net4(x, c_noise, cond):
            b = x.shape[0]                                   # the UNet reads the first B rows (no reference stream)
            eps, _ = O.unet_forward(sd, cfg, x, c_noise, cond["crossattn"][:b], cond["vector"][:b], cams=cams3,
                                    choices=list(range(nv)), cache=cache)
            return eps, None, None, None

        lat_o = O.euler_edm_sample(lambda x, s, cc: den(net4, x, s, cc)[0], noise.clone(), c_o, uc_o, steps, rows=3,
                                   scale=7.5, scale_im=3.5)
        img_o = V.decode_first_stage(vae_sd, vae_cfg, lat_o, V.SDXL_SCALE_FACTOR)
    assert samples.shape == lat_o.shape == (N, 4, L, L) and image.shape == img_o.shape
    r_c = _record("e2e_c_crossattn", c["crossattn"], c_o["crossattn"])
    r_lat = _record("e2e_latent_after_3_guided_steps", samples, lat_o)
    r_img = _record("e2e_decoded_image", image, img_o)
    # conditioner 1.5e-2 (its own test); 3 steps of CFG 7.5 amplify the per-evaluation bf16 error ~(1 + scale) / step
    assert r_c <= 1.5e-2 and r_lat <= 8e-2 and r_img <= 1e-1, (r_c, r_lat, r_img)
