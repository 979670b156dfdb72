"""Generate tests/golden/*.pt from the reference's OWN modules (imported in place from
/root/reference via oracle/ref_harness.py).  Run in the build container:

    python tests/golden/make_golden.py

Reference commit 1a23f97.  Weights/inputs come from oracle.sgm_oracle.synthetic_* (numpy
MT19937, platform independent), so the fixtures only hold the reference's OUTPUTS.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle import sgm_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LATENT = 16
N_VIEWS = 8
STEPS = 4


def cfg_inputs(cfg, seed=0):
    inp = O.synthetic_inputs(cfg, LATENT, n_img=1, seed=seed, n_views=N_VIEWS)
    c = {"crossattn": inp["crossattn"], "vector": inp["vector"]}
    # force_uc_zero_embeddings on the text keys (sample.py:155-161); `vector` keeps the size embeds
    uc = {"crossattn": torch.zeros_like(inp["crossattn"]), "vector": inp["vector"].clone()}
    uc["vector"][:, : cfg["adm_in_channels"] // 2] = 0
    return inp, c, uc


def main():
    torch.manual_seed(0)
    ns = H.import_reference()
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=0, latent=LATENT, num_references=N_VIEWS + 1)
    model = H.build_reference_unet(ns, cfg, sd)
    gold = {"reference_commit": "1a23f97", "latent": LATENT, "n_views": N_VIEWS, "steps": STEPS}

    # --- schedule / embeddings ---
    disc = ns.discretizer.LegacyDDPMDiscretization()
    gold["sigmas_50"] = disc(50)
    gold["sigmas_1000_flip"] = disc(1000, do_append_zero=False, flip=True)
    gold["sigmas_%d" % STEPS] = disc(STEPS)
    tt = torch.tensor([0.0, 1.0, 500.0, 999.0])
    gold["t_emb_in"] = tt
    gold["t_emb_320"] = ns.util.timestep_embedding(tt, 320)

    # --- one UNet evaluation, CFG batch of 3, pose conditioning on; then a cached second call ---
    inp, c, uc = cfg_inputs(cfg)
    cams = inp["cams"][0]
    pose = [H.cameras_from_packed(cams)] * 3
    choices = list(range(N_VIEWS))
    ns.sample.choices = choices
    x3 = torch.cat([inp["x"]] * 3)
    ctx3 = torch.cat([uc["crossattn"], uc["crossattn"], c["crossattn"]])
    y3 = torch.cat([uc["vector"], uc["vector"], c["vector"]])
    t3 = torch.tensor([500, 500, 500])
    with torch.no_grad():
        eps, fg, alphas, rgb = model(x3, timesteps=t3, context=ctx3, y=y3, pose=pose, mask_ref=None,
                                     drop_im=None)
        eps2, *_ = model(0.9 * x3, timesteps=t3 - 100, context=ctx3, y=y3, pose=pose, mask_ref=None,
                         drop_im=None)
    gold["unet_eps_step0"] = eps.float()
    gold["unet_eps_cached"] = eps2.float()
    gold["fg_masks"] = [f.float() for f in fg]
    gold["alphas"] = [a.float() for a in alphas]
    gold["rgbs"] = [r.float() for r in rgb]
    for m in model.modules():  # DiffusionEngine.clear_rendered_feat (diffusion.py:165-169)
        if hasattr(m, "pose_emb_layers"):
            m.rendered_feat = None

    # --- pose OFF plumbing config (BASELINE config 1 shape, tiny width) ---
    cfg_off = dict(cfg, image_cross_blocks=[])
    sd_off = O.synthetic_state_dict(cfg_off, seed=1)
    model_off = H.build_reference_unet(ns, cfg_off, sd_off)
    with torch.no_grad():
        eps_off, *_ = model_off(inp["x"], timesteps=torch.tensor([500]), context=c["crossattn"],
                                y=c["vector"])
    gold["unet_eps_pose_off"] = eps_off.float()

    # --- the reference's own sampler + denoiser + guider around the patched UNet, STEPS steps ---
    import importlib
    sampling = importlib.import_module("sgm.modules.diffusionmodules.sampling")
    denoiser_mod = importlib.import_module("sgm.modules.diffusionmodules.denoiser")
    wrappers = importlib.import_module("sgm.modules.diffusionmodules.wrappers")
    P = "sgm.modules.diffusionmodules."
    sampler = sampling.EulerEDMSampler(
        num_steps=STEPS, device="cpu",
        discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"},
        guider_config={"target": P + "guiders.ScheduledCFGImgTextRef",
                       "params": {"scale": 7.5, "scale_im": 3.5}})
    denoiser = denoiser_mod.DiscreteDenoiser(
        weighting_config={"target": P + "denoiser_weighting.EpsWeighting"},
        scaling_config={"target": P + "denoiser_scaling.EpsScaling"},
        num_idx=1000, discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"})
    net = wrappers.OpenAIWrapper(model)
    kwargs = {"pose": pose, "mask_ref": None, "drop_im": None}
    den = lambda inp_, sigma, cc: denoiser(net, inp_, sigma, cc, **kwargs)  # diffusion.py:392-394
    with torch.no_grad():
        samples, _ = sampler(den, inp["x"].clone(), c, uc=uc, num_steps=STEPS)
    gold["sample_final"] = samples.float()

    torch.save(gold, os.path.join(OUT, "tiny_unet_golden.pt"))
    print("wrote", os.path.join(OUT, "tiny_unet_golden.pt"))
    for k, v in gold.items():
        if isinstance(v, torch.Tensor):
            print(f"  {k}: {tuple(v.shape)} mean|.|={v.abs().mean():.4f}")


if __name__ == "__main__":
    main()
