"""Conditioner oracle (oracle/conditioner_oracle.py, groundwork for SURVEY §8f row 3) pinned against
the reference's own classes imported in place (GeneralConditioner, ConcatTimestepEmbedderND,
FrozenOpenCLIPEmbedder.encode_with_transformer), against Hugging Face's CLIPTextModel (the library
FrozenCLIPEmbedder calls) and against committed golden vectors (tests/golden/make_conditioner_golden.py)."""
import os
import sys
import types
from collections import OrderedDict

import pytest
import torch
import torch.nn as nn

from oracle import conditioner_oracle as C
from oracle import ref_harness as H

GOLD = os.path.join(os.path.dirname(__file__), "golden", "conditioner_golden.pt")
needs_ref = pytest.mark.skipif(not H.available(), reason="reference checkout not present")


def import_reference_encoders():
    """sgm/modules/encoders/modules.py imported in place: kornia / open_clip are not installed and are
    stood in for by empty modules (nothing on the text path touches them at import time)."""
    import importlib
    H.install()
    for name in ("kornia", "open_clip", "open_clip.tokenizer"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["open_clip.tokenizer"].SimpleTokenizer = object
    sys.modules["open_clip"].tokenizer = sys.modules["open_clip.tokenizer"]
    for sub in ("encoders", "autoencoding", "distributions"):
        name = "sgm.modules." + sub
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(H.REF, "sgm", "modules", sub)]
            sys.modules[name] = m
    return importlib.import_module("sgm.modules.encoders.modules")


def toy_batch(b=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"txt": r(b, 5), "txt_ref": r(3 * b, 5),
            "original_size_as_tuple": torch.tensor([[512.0, 512.0]]).repeat(b, 1),
            "original_size_as_tuple_ref": torch.tensor([[512.0, 384.0]]).repeat(3 * b, 1),
            "crop_coords_top_left": torch.tensor([[0.0, 16.0]]).repeat(b, 1),
            "crop_coords_top_left_ref": torch.tensor([[8.0, 0.0]]).repeat(3 * b, 1)}


# deterministic stand-ins for the two text encoders: [b, 5] -> [b, 4, dim] (and a pooled [b, dim])
def toy_text(x, dim, mul):
    seq = torch.arange(4, dtype=torch.float32)[None, :, None]
    ch = torch.arange(dim, dtype=torch.float32)[None, None, :]
    return torch.sin(mul * x.sum(-1)[:, None, None] + seq + 0.1 * ch)


def oracle_embedders():
    size = lambda v: C.concat_timestep_embedder_nd(v, 8)
    return [dict(fn=lambda v: toy_text(v, 6, 1.0), input_keys=["txt", "txt_ref"]),
            dict(fn=lambda v: (toy_text(v, 10, 0.5), toy_text(v, 7, 0.25)[:, 0]), input_keys=["txt", "txt_ref"]),
            dict(fn=size, input_keys=["original_size_as_tuple", "original_size_as_tuple_ref"]),
            dict(fn=size, input_keys=["crop_coords_top_left", "crop_coords_top_left_ref"])]


def reference_conditioner(mod):
    toy = types.ModuleType("cd360_toy_embedders")

    class ToyText(mod.AbstractEmbModel):
        def __init__(self, dim, mul, pooled_dim=0):
            super().__init__()
            self.dim, self.mul, self.pooled_dim, self.modifier_token = dim, mul, pooled_dim, None

        def forward(self, x):
            z = toy_text(x, self.dim, self.mul)
            return (z, toy_text(x, self.pooled_dim, 0.25)[:, 0]) if self.pooled_dim else z

    toy.ToyText = ToyText
    sys.modules["cd360_toy_embedders"] = toy
    size = {"target": "sgm.modules.encoders.modules.ConcatTimestepEmbedderND", "params": {"outdim": 8}}
    return mod.GeneralConditioner([
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": "cd360_toy_embedders.ToyText",
         "params": {"dim": 6, "mul": 1.0}},
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": "cd360_toy_embedders.ToyText",
         "params": {"dim": 10, "mul": 0.5, "pooled_dim": 7}},
        dict(size, is_trainable=False, input_keys="original_size_as_tuple,original_size_as_tuple_ref"),
        dict(size, is_trainable=False, input_keys="crop_coords_top_left,crop_coords_top_left_ref")])


class _Block(nn.Module):
    """open_clip ResidualAttentionBlock (published architecture): pre-LN, nn.MultiheadAttention, exact GELU."""

    def __init__(self, w, heads, mlp):
        super().__init__()
        self.ln_1 = nn.LayerNorm(w)
        self.attn = nn.MultiheadAttention(w, heads)
        self.ln_2 = nn.LayerNorm(w)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(w, mlp)), ("gelu", nn.GELU()), ("c_proj", nn.Linear(mlp, w))]))

    def forward(self, x, attn_mask=None):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


def open_clip_tower(cfg, sd):
    m = nn.Module()
    m.token_embedding = nn.Embedding(cfg["vocab"], cfg["width"])
    m.positional_embedding = nn.Parameter(torch.zeros(cfg["ctx"], cfg["width"]))
    m.transformer = nn.Module()
    m.transformer.resblocks = nn.ModuleList([_Block(cfg["width"], cfg["heads"], cfg["mlp"]) for _ in range(cfg["layers"])])
    m.transformer.grad_checkpointing = False
    m.ln_final = nn.LayerNorm(cfg["width"])
    m.text_projection = nn.Parameter(torch.zeros(cfg["width"], cfg["proj"]))
    m.register_buffer("attn_mask", C.causal_mask(cfg["ctx"]), persistent=False)
    missing, unexpected = m.load_state_dict({k[len("model."):]: v for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    return m.eval()


def tokens_for(cfg, b=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(1, cfg["vocab"] - 2, (b, cfg["ctx"]), generator=g)
    for i in range(b):                      # an eot token (highest id) somewhere, padding after it
        e = 3 + (4 * i) % (cfg["ctx"] - 5)
        t[i, e] = cfg["vocab"] - 1
        t[i, e + 1:] = 0
    return t


# ---------------------------------------------------------------------------------------------
@needs_ref
def test_general_conditioner_vs_reference():
    mod = import_reference_encoders()
    ref = reference_conditioner(mod)
    emb = oracle_embedders()
    batch = toy_batch()
    keys = [e["input_keys"] for e in emb]
    for force_ref in (False, True):
        with torch.no_grad():
            r = ref(dict(batch), force_ref_zero_embeddings=force_ref)
            rc, ruc = ref.get_unconditional_conditioning(dict(batch), force_uc_zero_embeddings=keys,
                                                         force_ref_zero_embeddings=force_ref)
        o = C.general_conditioner(emb, batch, None, force_ref)
        oc, ouc = C.get_unconditional_conditioning(emb, batch, None, keys, force_ref)
        assert set(r) == set(o) == {"crossattn", "vector"}
        rows = 2 if force_ref else 2 + 6
        assert o["crossattn"].shape == (rows, 4, 16) and o["vector"].shape == (rows, 7 + 16 + 16)
        for k in r:
            assert torch.equal(r[k], o[k]) and torch.equal(rc[k], oc[k]) and torch.equal(ruc[k], ouc[k]), (k, force_ref)
        assert float(ouc["crossattn"].abs().max()) == 0.0 and float(ouc["vector"].abs().max()) == 0.0   # sample.py:151-160
    # only the listed key pairs are zeroed
    part = C.general_conditioner(emb, batch, [["txt", "txt_ref"]], True)
    with torch.no_grad():
        rpart = ref(dict(batch), [["txt", "txt_ref"]], True)
    assert torch.equal(part["vector"], rpart["vector"]) and float(part["crossattn"].abs().max()) == 0.0
    assert float(part["vector"][:, :7].abs().max()) == 0.0 and float(part["vector"][:, 7:].abs().max()) > 0.0


@needs_ref
def test_concat_timestep_embedder_vs_reference():
    mod = import_reference_encoders()
    ref = mod.ConcatTimestepEmbedderND(256)
    x = torch.tensor([[512.0, 512.0], [1024.0, 768.0], [0.0, 33.0]])
    assert torch.equal(ref(x), C.concat_timestep_embedder_nd(x, 256))
    assert C.concat_timestep_embedder_nd(x, 256).shape == (3, 512)
    assert torch.equal(ref(x[:, 0]), C.concat_timestep_embedder_nd(x[:, 0], 256))


@needs_ref
def test_open_clip_text_vs_reference_control_flow():
    """The reference's own encode_with_transformer / text_transformer_forward / pool (penultimate =
    input of the last block, un-normalised; pooled at argmax of the ids) over a stand-in tower."""
    mod = import_reference_encoders()
    cfg = dict(C.TINY_OPEN_CLIP_CFG)
    sd = C.synthetic_state_dict(C.open_clip_param_shapes(cfg), seed=3)
    emb = mod.FrozenOpenCLIPEmbedder.__new__(mod.FrozenOpenCLIPEmbedder)
    nn.Module.__init__(emb)
    emb.model, emb.modifier_token, emb.legacy, emb.layer, emb.return_pooled = open_clip_tower(cfg, sd), None, False, "penultimate", True
    tok = tokens_for(cfg)
    with torch.no_grad():
        r = emb.encode_with_transformer(tok)
    o = C.open_clip_text(sd, cfg, tok)
    for k in ("penultimate", "last", "pooled"):
        assert float((r[k] - o[k]).abs().max()) <= 2e-5 * max(1.0, float(r[k].abs().max())), k
    assert o["penultimate"].shape == (3, cfg["ctx"], cfg["width"]) and o["pooled"].shape == (3, cfg["proj"])


def _hf_clip(cfg, sd):
    from transformers import CLIPTextConfig, CLIPTextModel
    hf = CLIPTextModel(CLIPTextConfig(vocab_size=cfg["vocab"], hidden_size=cfg["width"], intermediate_size=cfg["mlp"],
                                      num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                                      max_position_embeddings=cfg["ctx"], hidden_act="quick_gelu", layer_norm_eps=cfg["eps"],
                                      eos_token_id=cfg["vocab"] - 1, bos_token_id=1, pad_token_id=0)).eval()
    own = hf.state_dict()
    load = {k: sd["transformer." + k] for k in own if "transformer." + k in sd}
    missing = [k for k in own if k not in load and "position_ids" not in k]
    assert not missing, missing
    hf.load_state_dict(load, strict=False)
    return hf


def test_clip_text_hidden_vs_huggingface():
    """FrozenCLIPEmbedder.forward = embeddings -> causal encoder -> final_layer_norm of HF CLIPTextModel."""
    pytest.importorskip("transformers")
    cfg = dict(C.TINY_CLIP_CFG)
    sd = C.synthetic_state_dict(C.clip_param_shapes(cfg), seed=5)
    hf = _hf_clip(cfg, sd)
    tok = tokens_for(cfg)
    with torch.no_grad():
        ref = hf(input_ids=tok).last_hidden_state
    out = C.clip_text_hidden(sd, cfg, tok)
    assert out.shape == (3, cfg["ctx"], cfg["width"])
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    # causality: a change after position j leaves positions <= j untouched
    tok2 = tok.clone()
    tok2[:, 10:] = 5
    assert torch.equal(C.clip_text_hidden(sd, cfg, tok2)[:, :10], out[:, :10])


def test_sdxl_conditioner_shapes_and_golden():
    """The five embedders of the shipped yaml wired together (tiny towers): crossattn = clip | open_clip
    penultimate, vector = pooled | 3 x (2 x outdim); values against the committed golden vector."""
    clip_cfg, oc_cfg = dict(C.TINY_CLIP_CFG), dict(C.TINY_OPEN_CLIP_CFG)
    clip_sd = C.synthetic_state_dict(C.clip_param_shapes(clip_cfg), seed=5)
    oc_sd = C.synthetic_state_dict(C.open_clip_param_shapes(oc_cfg), seed=3)
    emb = C.sdxl_conditioner(clip_sd, clip_cfg, oc_sd, oc_cfg, size_dim=8)
    b = 2
    size = lambda v, n: torch.tensor([v]).repeat(n, 1)
    batch = {"txt": (tokens_for(clip_cfg, b, 1), tokens_for(oc_cfg, b, 2)),
             "txt_ref": (tokens_for(clip_cfg, 4 * b, 3), tokens_for(oc_cfg, 4 * b, 4)),
             "original_size_as_tuple": size([512.0, 512.0], b), "original_size_as_tuple_ref": size([512.0, 512.0], 4 * b),
             "crop_coords_top_left": size([0.0, 0.0], b), "crop_coords_top_left_ref": size([0.0, 0.0], 4 * b),
             "target_size_as_tuple": size([512.0, 512.0], b), "target_size_as_tuple_ref": size([512.0, 512.0], 4 * b)}
    c = C.general_conditioner(emb, batch)
    assert c["crossattn"].shape == (b + 4 * b, clip_cfg["ctx"], clip_cfg["width"] + oc_cfg["width"])
    assert c["vector"].shape == (b + 4 * b, oc_cfg["proj"] + 3 * 2 * 8)
    gold = torch.load(GOLD, weights_only=False)
    assert float((c["crossattn"] - gold["sdxl_tiny_crossattn"]).abs().max()) <= 1e-5
    assert float((c["vector"] - gold["sdxl_tiny_vector"]).abs().max()) <= 1e-5
    # golden outputs of the reference's own GeneralConditioner / HF CLIP (generated where both exist)
    o = C.general_conditioner(oracle_embedders(), toy_batch(), None, False)
    assert torch.equal(o["crossattn"], gold["toy_crossattn"]) and torch.equal(o["vector"], gold["toy_vector"])
    out = C.clip_text_hidden(clip_sd, clip_cfg, tokens_for(clip_cfg))
    assert float((out - gold["hf_clip_last_hidden"]).abs().max()) <= 2e-5 * float(gold["hf_clip_last_hidden"].abs().max())
