"""Generate tests/golden/sigma_idx_golden.pt from the reference's OWN DiscreteDenoiser /
LegacyDDPMDiscretization / CubicSampling / DiscreteSampling (imported in place from /root/reference
through oracle/ref_harness.py, reference commit 1a23f97).  Integer work: the σ → table-index
quantisation (denoiser.py:65-75), the 50 schedule indices (discretizer.py:11-14,42-69) and the
training-time σ draws (sigma_sampling.py:16-53) must match bit for bit.

    python tests/golden/make_sigma_golden.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sigma_idx_golden.pt")
P = "sgm.modules.diffusionmodules."


def main():
    H.import_reference()
    # sgm/util.py imports only stdlib + torch-side packages lazily: seed the shell first
    sys.modules["sgm"].util = importlib.import_module("sgm.util")
    den_mod = importlib.import_module(P + "denoiser")
    samp_mod = importlib.import_module(P + "sigma_sampling")
    den = den_mod.DiscreteDenoiser(
        weighting_config={"target": P + "denoiser_weighting.EpsWeighting"},
        scaling_config={"target": P + "denoiser_scaling.EpsScaling"}, num_idx=1000,
        discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"})
    disc = importlib.import_module(P + "discretizer").LegacyDDPMDiscretization()
    gold = {"reference_commit": "1a23f97", "table": den.sigmas.clone()}
    rng = np.random.RandomState(7)
    for n in (50, 36, 4):
        gold[f"sigmas_{n}"] = disc(n)
    tab = den.sigmas
    queries = torch.cat([
        disc(50), disc(36),
        (tab[:-1] + tab[1:]) / 2,                                  # exact midpoints: tie-breaking of argmin
        torch.tensor(np.exp(rng.uniform(np.log(0.02), np.log(20.0), 512)), dtype=torch.float32),
        torch.tensor([0.0, 1e-6, 14.6146, 14.61464, 100.0, 0.029167, 0.0291675]),
    ])
    gold["queries"] = queries
    gold["idx"] = den.sigma_to_idx(queries)
    gold["sigma_q"] = den.possibly_quantize_sigma(queries)
    c_skip, c_out, c_in, c_noise = den.scaling(gold["sigma_q"])
    gold["c_in"], gold["c_out"] = c_in, c_out
    gold["w"] = den.w(gold["sigma_q"])
    # training draws with injected uniforms (CubicSampling: rand ** 3; DiscreteSampling: randint)
    u = torch.tensor(rng.uniform(0, 1, 256), dtype=torch.float32)
    cubic = samp_mod.CubicSampling(discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"},
                                   num_idx=1000)
    gold["u"] = u
    t = (1 - u ** 3) * (cubic.num_idx - 1)
    gold["cubic_idx"] = t.long()
    gold["cubic_sigma"] = cubic.idx_to_sigma(t.long())
    torch.save(gold, OUT)
    print("wrote", OUT, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
