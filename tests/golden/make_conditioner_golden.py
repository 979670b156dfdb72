"""Generate tests/golden/conditioner_golden.pt: outputs of the reference's OWN GeneralConditioner
(sgm/modules/encoders/modules.py:73-230, imported in place) on the toy batch of
tests/test_oracle_conditioner.py, of Hugging Face's CLIPTextModel (the library FrozenCLIPEmbedder
calls) on a tiny seeded configuration, and — for regression only — the oracle's wired SDXL
conditioner on tiny towers.  Run in the build container:  python tests/golden/make_conditioner_golden.py
(reference commit 1a23f97)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import conditioner_oracle as C  # noqa: E402
from tests import test_oracle_conditioner as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    mod = T.import_reference_encoders()
    ref = T.reference_conditioner(mod)
    with torch.no_grad():
        r = ref(dict(T.toy_batch()), force_ref_zero_embeddings=False)
    clip_cfg, oc_cfg = dict(C.TINY_CLIP_CFG), dict(C.TINY_OPEN_CLIP_CFG)
    clip_sd = C.synthetic_state_dict(C.clip_param_shapes(clip_cfg), seed=5)
    oc_sd = C.synthetic_state_dict(C.open_clip_param_shapes(oc_cfg), seed=3)
    with torch.no_grad():
        hf_out = T._hf_clip(clip_cfg, clip_sd)(input_ids=T.tokens_for(clip_cfg)).last_hidden_state
    emb = C.sdxl_conditioner(clip_sd, clip_cfg, oc_sd, oc_cfg, size_dim=8)
    b = 2
    size = lambda v, n: torch.tensor([v]).repeat(n, 1)
    batch = {"txt": (T.tokens_for(clip_cfg, b, 1), T.tokens_for(oc_cfg, b, 2)),
             "txt_ref": (T.tokens_for(clip_cfg, 4 * b, 3), T.tokens_for(oc_cfg, 4 * b, 4)),
             "original_size_as_tuple": size([512.0, 512.0], b), "original_size_as_tuple_ref": size([512.0, 512.0], 4 * b),
             "crop_coords_top_left": size([0.0, 0.0], b), "crop_coords_top_left_ref": size([0.0, 0.0], 4 * b),
             "target_size_as_tuple": size([512.0, 512.0], b), "target_size_as_tuple_ref": size([512.0, 512.0], 4 * b)}
    c = C.general_conditioner(emb, batch)
    torch.save({"reference_commit": "1a23f97", "toy_crossattn": r["crossattn"], "toy_vector": r["vector"],
                "hf_clip_last_hidden": hf_out, "sdxl_tiny_crossattn": c["crossattn"], "sdxl_tiny_vector": c["vector"]},
               os.path.join(OUT, "conditioner_golden.pt"))
    print({k: tuple(v.shape) for k, v in r.items()}, tuple(hf_out.shape))


if __name__ == "__main__":
    main()
