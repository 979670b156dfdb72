"""Text / size conditioner of the reference (sgm/modules/encoders/modules.py:29-230, 377-517,
622-772, 1117-1134) on the sm_100a kernels — SURVEY.md §8f row 3.

Same class names, constructor kwargs (configs/train_co3d_concept.yaml:56-96), module tree and
state-dict keys as the reference, so `sd_xl_base_1.0.safetensors`' `conditioner.embedders.{0,1}.*`
entries load by name:
  * `FrozenCLIPEmbedder`      keys `transformer.text_model.…` (Hugging Face CLIPTextModel, CLIP-L)
  * `FrozenOpenCLIPEmbedder`  keys `model.…` (open_clip ViT-bigG-14 text tower); `FrozenOpenCLIPEmbedder2`
                              is the same tower without modifier tokens
  * `ConcatTimestepEmbedderND`, `GeneralConditioner`, `AbstractEmbModel`.

Arithmetic: embedding gather, causal attention, pooling gather in csrc/conditioner.cu; every
projection / MLP on the tcgen05 GEMM (quick-GELU / GELU in the epilogue, residual adds in place);
LayerNorm and the pooled projection on the shared row kernels.  bf16 activations, fp32 statistics.

Tokenisers are vocabulary FILES (BPE merges) that this image cannot download: `forward` accepts
token ids (int tensor [B, 77], what `CLIPTokenizer` / `open_clip.tokenize` produce) directly, or text
when a tokenizer object was attached (`embedder.tokenizer = …`; the reference builds it in __init__).

Bug-compat facts kept from the reference (oracle/conditioner_oracle.py pins them):
  * `layer: hidden, layer_idx: 11` of the shipped yaml is accepted and never read — FrozenCLIPEmbedder
    returns final_layer_norm(last layer) (modules.py:455-512);
  * "penultimate" = the INPUT of the last residual block, not layer-normed (:748-749);
  * pooled = ln_final(last)[argmax of the token ids] @ text_projection (:737-743).
Modifier tokens (`<new1>`, :418-431, 676-690): one extra row per token appended to the token
embedding, initialised from row 42170 ("ktn"; 47629 / 43514 for a second / third token); these rows
are what `main.py:611-625` saves as `embed` — `GeneralConditioner.modifier_token_rows()`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from .... import ops
from ...._lib import ACT_GELU, ACT_QUICK_GELU
from ...util import instantiate_from_config

bf16 = torch.bfloat16
MODIFIER_INIT_ROWS = (42170, 47629, 43514)


class AbstractEmbModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.is_trainable = None
        self.ucg_rate = None
        self.input_key = None
        self.modifier_token = None


def _pick(text, which: int):
    """A batch value that carries the ids of BOTH tokenisers, (clip_ids, open_clip_ids): the shipped
    yaml feeds the same `txt` key to both text embedders, each of which tokenises the string itself."""
    if isinstance(text, (tuple, list)) and len(text) == 2 and torch.is_tensor(text[0]):
        return text[which]
    return text


def _ids(tokens, device) -> torch.Tensor:
    if not torch.is_tensor(tokens):
        raise TypeError("expected token ids (int tensor [B, ctx]); attach a tokenizer to embed text")
    return tokens.to(device=device, dtype=torch.int32).contiguous()


class _Tower(nn.Module):
    """Shared runner of both text towers: pre-LN causal transformer over [B*ctx, w] bf16 tokens."""

    def _packs(self):
        dev = self._device()
        p = self.__dict__.get("_pk")
        if p is None or p["dev"] != dev:
            p = dict(dev=dev, layers=[self._pack_layer(i) for i in range(self.n_layers)])
            self.__dict__["_pk"] = p
        return p

    def invalidate_packed(self):
        self.__dict__["_pk"] = None

    def _run_layers(self, x, batch, ctx, act, upto=None, tap_last_input=False):
        """x bf16 [batch*ctx, w], updated in place.  Returns (x, input of the last block | None)."""
        tap = None
        layers = self._packs()["layers"]
        for i, L in enumerate(layers[:upto]):
            if tap_last_input and i == len(layers) - 1:
                tap = x.clone()
            h = ops.layernorm(x, L["g1"], L["b1"], eps=self.eps)
            qkv = ops.gemm(h, L["wqkv"], bias=L["bqkv"])
            w = self.width
            a = ops.attention_causal(qkv[:, :w], qkv[:, w:2 * w], qkv[:, 2 * w:], batch, self.heads, ctx)
            x = ops.gemm(a, L["wo"], bias=L["bo"], residual=x, out=x)
            h = ops.layernorm(x, L["g2"], L["b2"], eps=self.eps)
            h = ops.gemm(h, L["w1"], bias=L["bb1"], act=act)
            x = ops.gemm(h, L["w2"], bias=L["bb2"], residual=x, out=x)
        return x, tap


def _f(t):
    return t.detach().float().contiguous()


def _h(t):
    return t.detach().to(bf16).contiguous()


# ------------------------------------------------------------------------------------------------
# FrozenCLIPEmbedder — Hugging Face CLIPTextModel layout
# ------------------------------------------------------------------------------------------------
class _CLIPAttention(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(w, w) for _ in range(4))


class _CLIPMLP(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(w, m), nn.Linear(m, w)


class _CLIPLayer(nn.Module):
    def __init__(self, w, m, eps):
        super().__init__()
        self.self_attn = _CLIPAttention(w)
        self.layer_norm1 = nn.LayerNorm(w, eps=eps)
        self.mlp = _CLIPMLP(w, m)
        self.layer_norm2 = nn.LayerNorm(w, eps=eps)


class _CLIPEmbeddings(nn.Module):
    def __init__(self, vocab, w, ctx):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, w)
        self.position_embedding = nn.Embedding(ctx, w)
        self.register_buffer("position_ids", torch.arange(ctx).expand((1, -1)), persistent=False)


class _CLIPEncoder(nn.Module):
    def __init__(self, w, m, layers, eps):
        super().__init__()
        self.layers = nn.ModuleList([_CLIPLayer(w, m, eps) for _ in range(layers)])


class _CLIPTextTransformer(nn.Module):
    def __init__(self, vocab, w, m, layers, ctx, eps):
        super().__init__()
        self.embeddings = _CLIPEmbeddings(vocab, w, ctx)
        self.encoder = _CLIPEncoder(w, m, layers, eps)
        self.final_layer_norm = nn.LayerNorm(w, eps=eps)


class _CLIPTextModel(nn.Module):
    def __init__(self, vocab, w, m, layers, ctx, eps):
        super().__init__()
        self.text_model = _CLIPTextTransformer(vocab, w, m, layers, ctx, eps)

    def get_input_embeddings(self):
        return self.text_model.embeddings.token_embedding


CLIP_L_ARCH = dict(vocab=49408, width=768, heads=12, layers=12, mlp=3072, ctx=77, eps=1e-5)
OPEN_CLIP_ARCHS = {
    "ViT-bigG-14": dict(vocab=49408, width=1280, heads=20, layers=32, mlp=5120, ctx=77, eps=1e-5, proj=1280),
    "ViT-H-14": dict(vocab=49408, width=1024, heads=16, layers=24, mlp=4096, ctx=77, eps=1e-5, proj=1024),
}


def _split_modifier(tok) -> Optional[List[str]]:
    if tok is None:
        return None
    return tok.split("+") if "+" in tok else [tok]


def _append_modifier_rows(emb: nn.Embedding, n_new: int) -> nn.Embedding:
    """add_token (modules.py:418-431 / 676-690): `n_new` rows appended; the LAST new row copies row
    42170, the one before 47629, then 43514 (row ids taken modulo the vocabulary for toy sizes)."""
    old = emb.weight.data
    v = old.shape[0]
    new = nn.Embedding(v + n_new, old.shape[1], device=old.device, dtype=old.dtype)
    new.weight.data[:v] = old
    for k in range(n_new):
        new.weight.data[v + n_new - 1 - k] = old[MODIFIER_INIT_ROWS[k] % v]
    return new


class FrozenCLIPEmbedder(AbstractEmbModel, _Tower):
    """CLIP-L text encoder (reference :377-517).  `arch` (not in the reference) overrides the tower
    sizes for tests; the default is openai/clip-vit-large-patch14."""

    LAYERS = ["last", "pooled", "hidden"]

    def __init__(self, modifier_token=None, version="openai/clip-vit-large-patch14", device="cuda", max_length=77,
                 freeze=True, layer="last", layer_idx=None, always_return_pooled=False, arch: Optional[dict] = None):
        super().__init__()
        assert layer in self.LAYERS
        a = dict(CLIP_L_ARCH if arch is None else arch)
        if a["width"] != 64 * a["heads"]:
            raise NotImplementedError("the causal attention kernel is specialised for head dim 64")
        self.arch = a
        self.width, self.heads, self.n_layers, self.eps = a["width"], a["heads"], a["layers"], a["eps"]
        self.transformer = _CLIPTextModel(a["vocab"], a["width"], a["mlp"], a["layers"], a["ctx"], a["eps"])
        self.tokenizer = None
        self.device = device
        self.max_length = max_length
        self.modifier_token = _split_modifier(modifier_token)
        self.modifier_token_id: List[int] = []
        if self.modifier_token is not None:
            emb = self.transformer.text_model.embeddings
            n_new = len(self.modifier_token)
            self.modifier_token_id = list(range(a["vocab"], a["vocab"] + n_new))
            emb.token_embedding = _append_modifier_rows(emb.token_embedding, n_new)
        if freeze:
            self.freeze()
        self.layer = layer
        self.layer_idx = layer_idx          # accepted, never read (reference forward ignores it)
        self.return_pooled = always_return_pooled
        if layer == "hidden":
            assert layer_idx is not None and 0 <= abs(layer_idx) <= 12

    def freeze(self):
        self.transformer = self.transformer.eval()
        for p in self.parameters():
            p.requires_grad = False
        if self.modifier_token is not None:    # only the token embedding trains (:433-445)
            for p in self.transformer.get_input_embeddings().parameters():
                p.requires_grad = True

    def _device(self):
        return self.transformer.text_model.final_layer_norm.weight.device

    def _pack_layer(self, i):
        L = self.transformer.text_model.encoder.layers[i]
        a = L.self_attn
        return dict(g1=_f(L.layer_norm1.weight), b1=_f(L.layer_norm1.bias), g2=_f(L.layer_norm2.weight),
                    b2=_f(L.layer_norm2.bias),
                    wqkv=_h(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                    bqkv=_f(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                    wo=_h(a.out_proj.weight), bo=_f(a.out_proj.bias), w1=_h(L.mlp.fc1.weight), bb1=_f(L.mlp.fc1.bias),
                    w2=_h(L.mlp.fc2.weight), bb2=_f(L.mlp.fc2.bias))

    def tokenize(self, text):
        if self.tokenizer is None:
            raise RuntimeError("FrozenCLIPEmbedder: no tokenizer attached (vocabulary files are not part of this "
                               "build); pass token ids or set `.tokenizer` to a CLIPTokenizer")
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return enc["input_ids"]

    @torch.no_grad()
    def forward(self, text):
        text = _pick(text, 0)
        tokens = text if torch.is_tensor(text) else self.tokenize(text)
        tm = self.transformer.text_model
        dev = self._device()
        ids = _ids(tokens.view(-1, tokens.shape[-1]), dev)
        b, ctx = ids.shape
        x = ops.embed_tokens(ids, _f(tm.embeddings.token_embedding.weight), _f(tm.embeddings.position_embedding.weight))
        x, _ = self._run_layers(x, b, ctx, ACT_QUICK_GELU)
        z = ops.layernorm(x, _f(tm.final_layer_norm.weight), _f(tm.final_layer_norm.bias), eps=self.eps)
        return ops.cast_f32(z).view(b, ctx, self.width)

    def encode(self, text):
        return self(text)


# ------------------------------------------------------------------------------------------------
# FrozenOpenCLIPEmbedder — open_clip text tower layout
# ------------------------------------------------------------------------------------------------
class _OCAttention(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * w, w))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * w))
        self.out_proj = nn.Linear(w, w)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _OCMLP(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.c_fc, self.c_proj = nn.Linear(w, m), nn.Linear(m, w)


class _OCBlock(nn.Module):
    def __init__(self, w, m, eps):
        super().__init__()
        self.ln_1 = nn.LayerNorm(w, eps=eps)
        self.attn = _OCAttention(w)
        self.ln_2 = nn.LayerNorm(w, eps=eps)
        self.mlp = _OCMLP(w, m)


class _OCTransformer(nn.Module):
    def __init__(self, w, m, layers, eps):
        super().__init__()
        self.resblocks = nn.ModuleList([_OCBlock(w, m, eps) for _ in range(layers)])
        self.grad_checkpointing = False


class _OCTextTower(nn.Module):
    def __init__(self, a):
        super().__init__()
        w = a["width"]
        self.token_embedding = nn.Embedding(a["vocab"], w)
        self.positional_embedding = nn.Parameter(0.01 * torch.randn(a["ctx"], w))
        self.transformer = _OCTransformer(w, a["mlp"], a["layers"], a["eps"])
        self.ln_final = nn.LayerNorm(w, eps=a["eps"])
        self.text_projection = nn.Parameter(w ** -0.5 * torch.randn(w, a["proj"]))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6593)        # present in the checkpoint; unused here
        self.register_buffer("attn_mask", torch.full((a["ctx"], a["ctx"]), float("-inf")).triu_(1), persistent=False)


class FrozenOpenCLIPEmbedder(AbstractEmbModel, _Tower):
    """open_clip text tower (reference :622-772).  `arch` names an entry of OPEN_CLIP_ARCHS or is a
    dict of tower sizes (tests)."""

    LAYERS = ["last", "penultimate"]

    def __init__(self, modifier_token=None, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", max_length=77,
                 freeze=True, layer="last", always_return_pooled=False, legacy=True):
        super().__init__()
        assert layer in self.LAYERS
        a = dict(OPEN_CLIP_ARCHS[arch] if isinstance(arch, str) else arch)
        if a["width"] != 64 * a["heads"]:
            raise NotImplementedError("the causal attention kernel is specialised for head dim 64")
        self.arch = a
        self.width, self.heads, self.n_layers, self.eps = a["width"], a["heads"], a["layers"], a["eps"]
        self.model = _OCTextTower(a)
        self.tokenizer = None
        self.device = device
        self.max_length = max_length
        self.modifier_token = _split_modifier(modifier_token)
        self.modifier_token_id: List[int] = []
        self.return_pooled = always_return_pooled
        if self.modifier_token is not None:
            n_new = len(self.modifier_token)
            self.modifier_token_id = list(range(a["vocab"], a["vocab"] + n_new))
            self.model.token_embedding = _append_modifier_rows(self.model.token_embedding, n_new)
        if freeze:
            self.freeze()
        self.layer = layer
        self.layer_idx = 0 if layer == "last" else 1
        self.legacy = legacy

    def freeze(self):
        self.model = self.model.eval()
        for p in self.parameters():
            p.requires_grad = False
        if self.modifier_token is not None:    # (:692-705)
            for p in self.model.token_embedding.parameters():
                p.requires_grad = True

    def _device(self):
        return self.model.ln_final.weight.device

    def _pack_layer(self, i):
        B = self.model.transformer.resblocks[i]
        return dict(g1=_f(B.ln_1.weight), b1=_f(B.ln_1.bias), g2=_f(B.ln_2.weight), b2=_f(B.ln_2.bias),
                    wqkv=_h(B.attn.in_proj_weight), bqkv=_f(B.attn.in_proj_bias), wo=_h(B.attn.out_proj.weight),
                    bo=_f(B.attn.out_proj.bias), w1=_h(B.mlp.c_fc.weight), bb1=_f(B.mlp.c_fc.bias),
                    w2=_h(B.mlp.c_proj.weight), bb2=_f(B.mlp.c_proj.bias))

    def tokenize(self, texts, context_length=77):
        if self.tokenizer is None:
            raise RuntimeError("FrozenOpenCLIPEmbedder: no tokenizer attached (BPE vocabulary files are not part of "
                               "this build); pass token ids or set `.tokenizer`")
        return self.tokenizer(texts, context_length=context_length)

    @torch.no_grad()
    def forward(self, text):
        text = _pick(text, 1)
        tokens = text if torch.is_tensor(text) else self.tokenize(text)
        z = self.encode_with_transformer(tokens)
        if not self.return_pooled and self.legacy:
            return z
        if self.return_pooled:
            assert not self.legacy
            return z[self.layer], z["pooled"]
        return z[self.layer]

    def encode_with_transformer(self, text):
        m = self.model
        dev = self._device()
        ids = _ids(text, dev)
        b, ctx = ids.shape
        x = ops.embed_tokens(ids, _f(m.token_embedding.weight), _f(m.positional_embedding))
        want_pen = (not self.legacy) or self.layer == "penultimate"
        if self.legacy and self.layer == "penultimate":
            # legacy: ln_final of the penultimate state; the last block is never needed
            x, _ = self._run_layers(x, b, ctx, ACT_GELU, upto=self.n_layers - 1)
            z = ops.layernorm(x, _f(m.ln_final.weight), _f(m.ln_final.bias), eps=self.eps)
            return ops.cast_f32(z).view(b, ctx, self.width)
        x, pen = self._run_layers(x, b, ctx, ACT_GELU, tap_last_input=want_pen)
        o = ops.layernorm(x, _f(m.ln_final.weight), _f(m.ln_final.bias), eps=self.eps)
        if self.legacy:
            return ops.cast_f32(o).view(b, ctx, self.width)
        out = {"last": ops.cast_f32(x).view(b, ctx, self.width)}
        if pen is not None:
            out["penultimate"] = ops.cast_f32(pen).view(b, ctx, self.width)
        out["pooled"] = self.pool(o, text)
        return out

    def pool(self, x, text):
        """x: ln_final(last) as bf16 tokens [B*ctx, w] (or [B, ctx, w]); text: token ids [B, ctx].  The eot
        token has the highest id of each sequence (:737-743)."""
        b, ctx = text.shape
        tok = x.reshape(b * ctx, -1)
        if tok.dtype != bf16:
            tok = ops.cast_bf16(tok.float().contiguous())
        eot = text.to(tok.device).argmax(dim=-1).to(torch.int32) + torch.arange(b, device=tok.device, dtype=torch.int32) * ctx
        rows = ops.gather_rows(tok, eot.contiguous())
        wp = self.__dict__.get("_proj_pk")
        if wp is None or wp.device != tok.device:
            wp = _h(self.model.text_projection.t())
            self.__dict__["_proj_pk"] = wp
        return ops.small_linear(rows, wp)

    def text_transformer_forward(self, x, attn_mask=None):
        raise NotImplementedError("the tower runs fused inside encode_with_transformer (token layout)")

    def encode(self, text):
        return self(text)


class FrozenOpenCLIPEmbedder2(FrozenOpenCLIPEmbedder):
    """The variant without modifier tokens (reference :519-620)."""

    LAYERS = ["pooled", "last", "penultimate"]

    def __init__(self, arch="ViT-H-14", version="laion2b_s32b_b79k", device="cuda", max_length=77, freeze=True,
                 layer="last", always_return_pooled=False, legacy=True):
        super().__init__(None, arch, version, device, max_length, freeze, layer, always_return_pooled, legacy)


# ------------------------------------------------------------------------------------------------
# ConcatTimestepEmbedderND / GeneralConditioner
# ------------------------------------------------------------------------------------------------
class ConcatTimestepEmbedderND(AbstractEmbModel):
    """Every scalar of x [b, d] embedded on its own with the sinusoidal `Timestep(outdim)` embedding
    and concatenated: [b, d * outdim] (reference :1117-1134; openaimodel.Timestep -> util.timestep_embedding)."""

    def __init__(self, outdim):
        super().__init__()
        self.outdim = outdim

    @torch.no_grad()
    def forward(self, x):
        if x.ndim == 1:
            x = x[:, None]
        assert len(x.shape) == 2
        b, dims = x.shape
        emb = ops.timestep_embedding(x.reshape(-1).float().contiguous(), self.outdim)
        return emb.view(b, dims * self.outdim)


def disabled_train(self, mode=True):
    return self


class GeneralConditioner(nn.Module):
    OUTPUT_DIM2KEYS = {2: "vector", 3: "crossattn", 4: "concat", 5: "concat"}
    KEY2CATDIM = {"vector": 1, "crossattn": 2, "concat": 1}

    def __init__(self, emb_models):
        super().__init__()
        embedders = []
        for embconfig in emb_models:
            embedder = instantiate_from_config(embconfig)
            assert isinstance(embedder, AbstractEmbModel), \
                f"embedder model {embedder.__class__.__name__} has to inherit from AbstractEmbModel"
            embedder.is_trainable = embconfig.get("is_trainable", False)
            embedder.ucg_rate = embconfig.get("ucg_rate", 0.0)
            if embedder.ucg_rate:
                raise NotImplementedError("ucg_rate > 0 (conditioning dropout) is unused by the shipped config")
            if not embedder.is_trainable:
                embedder.train = disabled_train.__get__(embedder)
                embedder.eval()
            if "input_key" in embconfig:
                embedder.input_key = embconfig["input_key"]
            elif "input_keys" in embconfig:
                embedder.input_keys = embconfig["input_keys"].split(",")
            else:
                raise KeyError(f"need either 'input_key' or 'input_keys' for embedder {embedder.__class__.__name__}")
            if embconfig.get("legacy_ucg_value", None) is not None:
                raise NotImplementedError("legacy_ucg_value is unused by the shipped config")
            embedder.legacy_ucg_val = None
            embedders.append(embedder)
        self.embedders = nn.ModuleList(embedders)

    def forward(self, batch: Dict, force_zero_embeddings: Optional[List] = None,
                force_ref_zero_embeddings: bool = False) -> Dict:
        """Reference :122-208: per-embedder outputs grouped by rank (2 -> vector, 3 -> crossattn), split
        into the main / reference halves when an embedder was declared with `input_keys`, concatenated
        feature-wise across embedders; reference rows are appended to the batch axis at the end."""
        output: Dict[str, torch.Tensor] = {}
        force_zero_embeddings = force_zero_embeddings or []
        for embedder in self.embedders:
            keys = getattr(embedder, "input_keys", None)
            if embedder.input_key is not None:
                emb_out = embedder(batch[embedder.input_key])
            elif keys is not None:
                if force_ref_zero_embeddings:
                    emb_out = embedder(batch[keys[0]])
                else:
                    emb_out = [embedder(batch[k]) for k in keys]
                    if isinstance(emb_out[0], tuple):
                        emb_out = [torch.cat([x[0] for x in emb_out]), torch.cat([x[1] for x in emb_out])]
                    else:
                        emb_out = torch.cat(emb_out)
            else:
                raise KeyError("embedder has neither input_key nor input_keys")
            assert isinstance(emb_out, (torch.Tensor, list, tuple))
            if not isinstance(emb_out, (list, tuple)):
                emb_out = [emb_out]
            for emb in emb_out:
                out_key = self.OUTPUT_DIM2KEYS[emb.dim()]
                if embedder.input_key is not None and embedder.input_key in force_zero_embeddings:
                    emb = torch.zeros_like(emb)
                if keys is not None and keys in force_zero_embeddings:      # a LIST matched against the entries
                    emb = torch.zeros_like(emb)
                if out_key in output:
                    if keys is not None:
                        catdim = 1 if ("pose" in keys) else self.KEY2CATDIM[out_key]
                        if not force_ref_zero_embeddings:
                            c, c1 = emb.chunk(2)
                            output[out_key] = torch.cat((output[out_key], c), catdim)
                            output[out_key + "_ref"] = torch.cat((output[out_key + "_ref"], c1), catdim)
                        else:
                            output[out_key] = torch.cat((output[out_key], emb), catdim)
                    else:
                        catdim = 1 if ("pose" in embedder.input_key and emb.size(1) != 77) else self.KEY2CATDIM[out_key]
                        output[out_key] = torch.cat((output[out_key], emb), catdim)
                else:
                    if keys is not None and not force_ref_zero_embeddings:
                        output[out_key], output[out_key + "_ref"] = emb.chunk(2)
                    else:
                        output[out_key] = emb
        for out_key in self.OUTPUT_DIM2KEYS.values():
            if out_key + "_ref" in output and not force_ref_zero_embeddings:
                output[out_key] = torch.cat([output[out_key], output[out_key + "_ref"]], 0)
                del output[out_key + "_ref"]
        return output

    def get_unconditional_conditioning(self, batch_c, batch_uc=None, force_uc_zero_embeddings=None,
                                       force_ref_zero_embeddings=None):
        c = self(batch_c, force_ref_zero_embeddings=force_ref_zero_embeddings)
        uc = self(batch_c if batch_uc is None else batch_uc, force_uc_zero_embeddings or [], force_ref_zero_embeddings)
        return c, uc

    def modifier_token_rows(self) -> List[torch.Tensor]:
        """`embed` of the delta checkpoint (main.py:623-624): the last row of the token embedding of
        embedders 0 (CLIP-L) and 1 (OpenCLIP)."""
        e0, e1 = self.embedders[0], self.embedders[1]
        return [e0.transformer.text_model.embeddings.token_embedding.weight[-1:].detach().clone(),
                e1.model.token_embedding.weight[-1:].detach().clone()]

    def load_modifier_token_rows(self, embed):
        """Delta-checkpoint load (sgm/util.py:225-228): replace the appended rows by the trained ones."""
        e0, e1 = self.embedders[0], self.embedders[1]
        with torch.no_grad():
            e0.transformer.text_model.embeddings.token_embedding.weight[-1:].copy_(embed[0])
            e1.model.token_embedding.weight[-1:].copy_(embed[1])
