"""Generate tests/golden/vae_decoder_golden.pt from the reference's OWN `Decoder`
(sgm/modules/diffusionmodules/model.py:604-757, imported in place via oracle/ref_harness.py) and the
1x1 post_quant_conv of AutoencoderKL.decode (sgm/models/autoencoder.py:313-316).  Run in the build
container:  python tests/golden/make_vae_golden.py      (reference commit 1a23f97)

Weights / latents come from oracle.vae_oracle.synthetic_state_dict and a seeded CPU generator, so the
fixture only holds the reference's OUTPUT."""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from oracle import vae_oracle as V  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED_W, SEED_Z, LATENT, BATCH = 1, 0, 16, 2


def latent(seed=SEED_Z, batch=BATCH, size=LATENT):
    return torch.randn(batch, 4, size, size, generator=torch.Generator().manual_seed(seed))


def main():
    H.install()
    m = importlib.import_module("sgm.modules.diffusionmodules.model")
    cfg = dict(V.TINY_VAE_CFG)
    sd = V.synthetic_state_dict(cfg, seed=SEED_W)
    dec = m.Decoder(**cfg).eval()
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    z = latent()
    with torch.no_grad():
        zq = F.conv2d(z * (1.0 / V.SDXL_SCALE_FACTOR), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
        img = dec(zq)
    torch.save({"reference_commit": "1a23f97", "seed_w": SEED_W, "seed_z": SEED_Z, "latent": LATENT, "batch": BATCH,
                "scale_factor": V.SDXL_SCALE_FACTOR, "image": img.half()}, os.path.join(OUT, "vae_decoder_golden.pt"))
    print("image", tuple(img.shape), "std %.4f" % float(img.std()))


if __name__ == "__main__":
    main()
