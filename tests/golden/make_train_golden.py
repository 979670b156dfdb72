"""Generate tests/golden/train_step_golden.pt from the REFERENCE's own training code (run where
/root/reference exists): StandardDiffusionLossImgRef -> DiscreteDenoiser -> OpenAIWrapper ->
UNetModel(.train(), stratified jitter on) under torch.autograd, tiny config, seeded weights/batch.

Stored: the random draws the reference made (so the oracle / CUDA path can replay them), the loss
terms, and for every trainable ('pose') parameter gradient its L2 norm, sum and a strided sample of
256 values (the full gradients are ~13 MB; the sample + two moments pin them).

    python tests/golden/make_train_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import ref_harness as H  # noqa: E402
from oracle import sgm_oracle as O  # noqa: E402
from oracle import train_oracle as T  # noqa: E402
from tests.test_oracle_vs_reference import _reference_training_step  # noqa: E402

CASE = dict(latent=16, n_views=3, b=2, weights_seed=2, batch_seed=5, image=48, torch_seed=11, drop_im=[1.0, 1.0])


def sample(g: torch.Tensor, n: int = 256) -> torch.Tensor:
    flat = g.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].clone()


def main():
    ns = H.import_reference()
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=CASE["weights_seed"])
    batch = T.synthetic_train_batch(cfg, CASE["latent"], n_views=CASE["n_views"], b=CASE["b"], seed=CASE["batch_seed"],
                                    image=CASE["image"])
    batch["drop_im"] = torch.tensor(CASE["drop_im"])
    total, terms, grads, rand = _reference_training_step(ns, cfg, sd, batch, seed=CASE["torch_seed"], train_mode=True)
    out = dict(case=CASE, reference_commit="1a23f97", total=float(total), terms={k: float(v) for k, v in terms.items()},
               rand=rand, grads={k: dict(norm=float(g.norm()), sum=float(g.sum()), sample=sample(g)) for k, g in grads.items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_step_golden.pt")
    torch.save(out, path)
    print(f"wrote {path}: total {float(total):.6f}, {len(grads)} gradient tensors, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
