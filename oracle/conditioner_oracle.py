"""CPU oracle of the conditioner — TEST INFRASTRUCTURE ONLY, groundwork for SURVEY.md §8f row 3.

The product path for this row (text encoders on the sm_100a kernels) is NOT built yet; per the
scope order the oracle comes first.  Plain-torch fp32, functional (state dict in, tensors out), one
function per reference symbol of sgm/modules/encoders/modules.py:

  * `clip_text_hidden`            FrozenCLIPEmbedder.forward / custom_forward (:455-512): HF
                                   `CLIPTextModel` embeddings -> causal pre-LN encoder (quick_gelu MLP)
                                   -> final_layer_norm.  Bug-compat: `layer: hidden, layer_idx: 11` of
                                   the shipped yaml (:65-66) is accepted by the constructor but never
                                   read by forward — the LAST layer's normalised states are returned.
  * `open_clip_text`              FrozenOpenCLIPEmbedder.encode_with_transformer / text_transformer_forward
                                   / pool (:716-762): token + positional embedding, causal transformer
                                   of `ResidualAttentionBlock`s (exact GELU), "penultimate" = the INPUT
                                   of the last block (not layer-normed), pooled = ln_final(last)[eot]
                                   @ text_projection with eot = argmax of the token ids.
  * `concat_timestep_embedder_nd` ConcatTimestepEmbedderND.forward (:1126-1134) over `Timestep`
                                   (openaimodel.py: sinusoidal `timestep_embedding`).
  * `general_conditioner`         GeneralConditioner.forward (:122-208) incl. the `input_keys`
                                   (txt, txt_ref) pairing, chunk(2) into main / reference halves, the
                                   final batch concatenation and both `force_*zero*` switches;
                                   `get_unconditional_conditioning` (:210-230).
  * `sdxl_conditioner`            the five embedders of configs/train_co3d_concept.yaml:56-96 wired
                                   together: crossattn [.., 77, 768 + 1280], vector [.., 1280 + 6 * 256].

Parity pinning (tests/test_oracle_conditioner.py): `general_conditioner` and
`concat_timestep_embedder_nd` against the reference's own classes imported in place;
`clip_text_hidden` against Hugging Face `CLIPTextModel` (the library the reference calls; the installed
transformers is newer than the reference's pin, so the pin is on the library's current semantics);
`open_clip_text` drives the reference's own `encode_with_transformer` over a stand-in text tower built
from torch.nn.MultiheadAttention blocks — open_clip itself is not installed, so its block internals
are restated from the published architecture: that part is "parity unpinned".  Tokenisers (BPE
vocabularies) are outside the oracle: inputs are token ids.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

CLIP_L_CFG = dict(vocab=49408, width=768, heads=12, layers=12, mlp=3072, ctx=77, eps=1e-5)
OPEN_CLIP_BIGG_CFG = dict(vocab=49408, width=1280, heads=20, layers=32, mlp=5120, ctx=77, eps=1e-5, proj=1280)
TINY_CLIP_CFG = dict(vocab=96, width=64, heads=2, layers=3, mlp=128, ctx=16, eps=1e-5)
TINY_OPEN_CLIP_CFG = dict(vocab=96, width=64, heads=2, layers=3, mlp=128, ctx=16, eps=1e-5, proj=48)


def causal_mask(n: int) -> Tensor:
    """Additive mask, -inf above the diagonal (modules.py:447-453; open_clip `build_attention_mask`)."""
    return torch.full((n, n), float("-inf")).triu_(1)


def _mha(x: Tensor, wq, bq, wk, bk, wv, bv, wo, bo, heads: int, mask: Tensor) -> Tensor:
    b, n, c = x.shape
    d = c // heads
    q, k, v = (F.linear(x, w, bb).view(b, n, heads, d).transpose(1, 2) for w, bb in ((wq, bq), (wk, bk), (wv, bv)))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d) + mask, dim=-1)
    return F.linear((att @ v).transpose(1, 2).reshape(b, n, c), wo, bo)


def quick_gelu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(1.702 * x)


# --------------------------------------------------------------------------------------------
# FrozenCLIPEmbedder (HF CLIPTextModel, openai/clip-vit-large-patch14)
# --------------------------------------------------------------------------------------------
def clip_text_hidden(sd: Dict[str, Tensor], cfg: dict, tokens: Tensor, prefix: str = "transformer.text_model.") -> Tensor:
    """tokens int64 [B, ctx] -> [B, ctx, width]: final_layer_norm(encoder(embeddings(tokens)))."""
    g = lambda k: sd[prefix + k]
    n = tokens.shape[1]
    x = g("embeddings.token_embedding.weight")[tokens] + g("embeddings.position_embedding.weight")[:n]
    mask = causal_mask(n)
    for i in range(cfg["layers"]):
        p = f"encoder.layers.{i}."
        h = F.layer_norm(x, (cfg["width"],), g(p + "layer_norm1.weight"), g(p + "layer_norm1.bias"), cfg["eps"])
        x = x + _mha(h, g(p + "self_attn.q_proj.weight"), g(p + "self_attn.q_proj.bias"),
                     g(p + "self_attn.k_proj.weight"), g(p + "self_attn.k_proj.bias"),
                     g(p + "self_attn.v_proj.weight"), g(p + "self_attn.v_proj.bias"),
                     g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"), cfg["heads"], mask)
        h = F.layer_norm(x, (cfg["width"],), g(p + "layer_norm2.weight"), g(p + "layer_norm2.bias"), cfg["eps"])
        x = x + F.linear(quick_gelu(F.linear(h, g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias"))),
                         g(p + "mlp.fc2.weight"), g(p + "mlp.fc2.bias"))
    return F.layer_norm(x, (cfg["width"],), g("final_layer_norm.weight"), g("final_layer_norm.bias"), cfg["eps"])


def clip_param_shapes(cfg: dict, prefix: str = "transformer.text_model.") -> Dict[str, tuple]:
    w, m = cfg["width"], cfg["mlp"]
    out = {prefix + "embeddings.token_embedding.weight": (cfg["vocab"], w),
           prefix + "embeddings.position_embedding.weight": (cfg["ctx"], w),
           prefix + "final_layer_norm.weight": (w,), prefix + "final_layer_norm.bias": (w,)}
    for i in range(cfg["layers"]):
        p = f"{prefix}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out[p + f"self_attn.{n}.weight"] = (w, w)
            out[p + f"self_attn.{n}.bias"] = (w,)
        for n in ("layer_norm1", "layer_norm2"):
            out[p + n + ".weight"] = (w,)
            out[p + n + ".bias"] = (w,)
        out[p + "mlp.fc1.weight"], out[p + "mlp.fc1.bias"] = (m, w), (m,)
        out[p + "mlp.fc2.weight"], out[p + "mlp.fc2.bias"] = (w, m), (w,)
    return out


# --------------------------------------------------------------------------------------------
# FrozenOpenCLIPEmbedder (open_clip ViT-bigG-14 text tower)
# --------------------------------------------------------------------------------------------
def open_clip_text(sd: Dict[str, Tensor], cfg: dict, tokens: Tensor, prefix: str = "model.") -> Dict[str, Tensor]:
    """tokens int64 [B, ctx] -> {"penultimate": [B, ctx, w], "last": [B, ctx, w], "pooled": [B, proj]}."""
    g = lambda k: sd[prefix + k]
    w, n = cfg["width"], tokens.shape[1]
    x = g("token_embedding.weight")[tokens] + g("positional_embedding")[:n]
    mask = causal_mask(n)
    out = {}
    for i in range(cfg["layers"]):
        if i == cfg["layers"] - 1:
            out["penultimate"] = x                                     # modules.py:748-749
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (w,), g(p + "ln_1.weight"), g(p + "ln_1.bias"), cfg["eps"])
        wi, bi = g(p + "attn.in_proj_weight"), g(p + "attn.in_proj_bias")
        x = x + _mha(h, wi[:w], bi[:w], wi[w:2 * w], bi[w:2 * w], wi[2 * w:], bi[2 * w:],
                     g(p + "attn.out_proj.weight"), g(p + "attn.out_proj.bias"), cfg["heads"], mask)
        h = F.layer_norm(x, (w,), g(p + "ln_2.weight"), g(p + "ln_2.bias"), cfg["eps"])
        x = x + F.linear(F.gelu(F.linear(h, g(p + "mlp.c_fc.weight"), g(p + "mlp.c_fc.bias"))),
                         g(p + "mlp.c_proj.weight"), g(p + "mlp.c_proj.bias"))
    out["last"] = x
    o = F.layer_norm(x, (w,), g("ln_final.weight"), g("ln_final.bias"), cfg["eps"])
    # pool (:737-743): the eot token has the highest id of each sequence
    out["pooled"] = o[torch.arange(o.shape[0]), tokens.argmax(dim=-1)] @ g("text_projection")
    return out


def open_clip_param_shapes(cfg: dict, prefix: str = "model.") -> Dict[str, tuple]:
    w, m = cfg["width"], cfg["mlp"]
    out = {prefix + "token_embedding.weight": (cfg["vocab"], w), prefix + "positional_embedding": (cfg["ctx"], w),
           prefix + "ln_final.weight": (w,), prefix + "ln_final.bias": (w,), prefix + "text_projection": (w, cfg["proj"])}
    for i in range(cfg["layers"]):
        p = f"{prefix}transformer.resblocks.{i}."
        out[p + "attn.in_proj_weight"], out[p + "attn.in_proj_bias"] = (3 * w, w), (3 * w,)
        out[p + "attn.out_proj.weight"], out[p + "attn.out_proj.bias"] = (w, w), (w,)
        for n in ("ln_1", "ln_2"):
            out[p + n + ".weight"] = (w,)
            out[p + n + ".bias"] = (w,)
        out[p + "mlp.c_fc.weight"], out[p + "mlp.c_fc.bias"] = (m, w), (m,)
        out[p + "mlp.c_proj.weight"], out[p + "mlp.c_proj.bias"] = (w, m), (w,)
    return out


def synthetic_state_dict(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if len(shape) == 2:
            std = 0.02 if "embedding" in name else 1.0 / math.sqrt(shape[-1] if "text_projection" not in name else shape[0])
            sd[name] = std * torch.randn(shape, generator=g)
        elif name.endswith("weight"):                  # LayerNorm gains
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.02 * torch.randn(shape, generator=g)
    return sd


# --------------------------------------------------------------------------------------------
# ConcatTimestepEmbedderND
# --------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """diffusionmodules/util.py:206-230 (cos | sin), as `Timestep(dim)` applies it."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def concat_timestep_embedder_nd(x: Tensor, outdim: int) -> Tensor:
    """[b, d] (or [b]) -> [b, d * outdim]: every scalar embedded on its own, concatenated (:1126-1134)."""
    if x.ndim == 1:
        x = x[:, None]
    b, dims = x.shape
    return timestep_embedding(x.reshape(-1), outdim).reshape(b, dims * outdim)


# --------------------------------------------------------------------------------------------
# GeneralConditioner
# --------------------------------------------------------------------------------------------
OUTPUT_DIM2KEYS = {2: "vector", 3: "crossattn", 4: "concat", 5: "concat"}
KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}


def general_conditioner(embedders: Sequence[dict], batch: dict, force_zero_embeddings: Optional[list] = None,
                        force_ref_zero_embeddings: bool = False) -> Dict[str, Tensor]:
    """GeneralConditioner.forward (:122-208) for embedders declared with `input_keys` (the shipped
    config).  embedders: [{"fn": callable(batch value) -> tensor | (tensor, tensor), "input_keys": [k, k_ref]}].
    Each embedder's outputs are grouped by rank (2 -> "vector", 3 -> "crossattn"), split into the main
    and the reference half, concatenated feature-wise across embedders, and finally the reference
    rows are appended to the batch axis."""
    force_zero_embeddings = force_zero_embeddings or []
    output: Dict[str, Tensor] = {}
    for e in embedders:
        keys = e["input_keys"]
        if force_ref_zero_embeddings:
            emb_out = e["fn"](batch[keys[0]])
        else:
            emb_out = [e["fn"](batch[k]) for k in keys]
            if isinstance(emb_out[0], tuple):
                emb_out = [torch.cat([x[0] for x in emb_out]), torch.cat([x[1] for x in emb_out])]
            else:
                emb_out = torch.cat(emb_out)
        if not isinstance(emb_out, (list, tuple)):
            emb_out = [emb_out]
        for emb in emb_out:
            out_key = OUTPUT_DIM2KEYS[emb.dim()]
            if keys in force_zero_embeddings:              # a LIST compared with the entries (:170-173)
                emb = torch.zeros_like(emb)
            catdim = 1 if ("pose" in keys) else KEY2CATDIM[out_key]
            if out_key in output:
                if not force_ref_zero_embeddings:
                    c, c1 = emb.chunk(2)
                    output[out_key] = torch.cat((output[out_key], c), catdim)
                    output[out_key + "_ref"] = torch.cat((output[out_key + "_ref"], c1), catdim)
                else:
                    output[out_key] = torch.cat((output[out_key], emb), catdim)
            else:
                if not force_ref_zero_embeddings:
                    output[out_key], output[out_key + "_ref"] = emb.chunk(2)
                else:
                    output[out_key] = emb
    for out_key in OUTPUT_DIM2KEYS.values():
        if out_key + "_ref" in output and not force_ref_zero_embeddings:
            output[out_key] = torch.cat([output[out_key], output[out_key + "_ref"]], 0)
            del output[out_key + "_ref"]
    return output


def get_unconditional_conditioning(embedders, batch_c, batch_uc=None, force_uc_zero_embeddings=None,
                                   force_ref_zero_embeddings=False):
    """(:210-230).  sample.py:151-160 passes every embedder's `input_keys` list as
    force_uc_zero_embeddings, so `uc` is all zeros (crossattn AND vector), and
    force_ref_zero_embeddings=True (only the first key of each pair is embedded)."""
    c = general_conditioner(embedders, batch_c, None, force_ref_zero_embeddings)
    uc = general_conditioner(embedders, batch_c if batch_uc is None else batch_uc,
                             force_uc_zero_embeddings or [], force_ref_zero_embeddings)
    return c, uc


def sdxl_conditioner(clip_sd, clip_cfg, oc_sd, oc_cfg, size_dim: int = 256) -> List[dict]:
    """The five embedders of configs/train_co3d_concept.yaml:56-96 over TOKEN IDS: batch["txt"] /
    batch["txt_ref"] = (clip_tokens, open_clip_tokens) pairs."""
    def clip(v):
        return clip_text_hidden(clip_sd, clip_cfg, v[0])

    def oc(v):
        o = open_clip_text(oc_sd, oc_cfg, v[1])
        return o["penultimate"], o["pooled"]               # layer: penultimate, always_return_pooled (:711-713)

    size = lambda v: concat_timestep_embedder_nd(v, size_dim)
    return [dict(fn=clip, input_keys=["txt", "txt_ref"]), dict(fn=oc, input_keys=["txt", "txt_ref"]),
            dict(fn=size, input_keys=["original_size_as_tuple", "original_size_as_tuple_ref"]),
            dict(fn=size, input_keys=["crop_coords_top_left", "crop_coords_top_left_ref"]),
            dict(fn=size, input_keys=["target_size_as_tuple", "target_size_as_tuple_ref"])]
