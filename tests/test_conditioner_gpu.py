"""Product conditioner (custom_diffusion360_b200/sgm/modules/encoders/modules.py)
on the sm_100a kernels against oracle/conditioner_oracle.py — the restatement pinned to the
reference's own GeneralConditioner / encode_with_transformer and to Hugging Face's CLIPTextModel
(tests/test_oracle_conditioner.py).

Tolerances: bf16 residual stream (2^-9 per rounding), 3 sequential roundings per layer, pre-LN
transformer -> rel_rms <= 1.5e-2 for the 3-layer toy towers, <= 3e-2 for the 32-layer bigG tower;
integer work (eot index, token ids) exact.  Metrics -> gpurun_out/conditioner_parity_metrics.json."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import conditioner_oracle as C

gpu = pytest.mark.gpu
METRICS = {}
E = "custom_diffusion360_b200.sgm.modules.encoders.modules."
# toy towers with the head width the kernels are built for (64)
TINY_CLIP = dict(vocab=96, width=128, heads=2, layers=3, mlp=256, ctx=16, eps=1e-5)
TINY_OC = dict(vocab=96, width=128, heads=2, layers=3, mlp=256, ctx=16, eps=1e-5, proj=48)


def _check(name, ours, ref, rel_tol):
    ours, ref = ours.detach().float().cpu(), ref.detach().float().cpu()
    rel = float((ours - ref).norm() / ref.norm().clamp_min(1e-12))
    mx = float((ours - ref).abs().max())
    METRICS[name] = dict(rel_rms=rel, max_abs=mx, ref_max=float(ref.abs().max()))
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "conditioner_parity_metrics.json"), "w") as f:
        json.dump(METRICS, f, indent=1)
    assert rel <= rel_tol, f"{name}: rel_rms {rel:.4g} > {rel_tol}"


def _tokens(cfg, b, seed, eot=True):
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(1, cfg["vocab"] - 2, (b, cfg["ctx"]), generator=g)
    if eot:   # the end-of-text token carries the highest id; padding after it repeats smaller ids
        pos = torch.randint(2, cfg["ctx"], (b,), generator=g)
        for i in range(b):
            t[i, pos[i]] = cfg["vocab"] - 1
            t[i, pos[i] + 1:] = 0
    return t


@gpu
def test_causal_attention_kernel():
    from custom_diffusion360_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for b, h, n in ((2, 12, 77), (3, 20, 77), (1, 2, 16), (2, 4, 128), (1, 1, 1)):
        w = h * 64
        qkv = torch.randn(b * n, 3 * w, generator=g).to(torch.bfloat16)
        out = ops.attention_causal(qkv[:, :w].to(dev), qkv[:, w:2 * w].to(dev), qkv[:, 2 * w:].to(dev), b, h, n)
        q, k, v = (qkv[:, i * w:(i + 1) * w].float().view(b, n, h, 64).transpose(1, 2) for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b * n, w)
        _check(f"attention_causal_b{b}_h{h}_n{n}", out, ref, 1.5 * 2 ** -9)
    # strided views of one fused buffer (what the towers pass)
    b, h, n = 2, 2, 16
    w = h * 64
    qkv = torch.randn(b * n, 3 * w, generator=g).to(torch.bfloat16).to(dev)
    out = ops.attention_causal(qkv[:, :w], qkv[:, w:2 * w], qkv[:, 2 * w:], b, h, n)
    q, k, v = (qkv[:, i * w:(i + 1) * w].float().cpu().view(b, n, h, 64).transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b * n, w)
    _check("attention_causal_fused_views", out, ref, 1.5 * 2 ** -9)


@gpu
def test_embed_gather_and_activation_epilogues():
    from custom_diffusion360_b200 import ops
    from custom_diffusion360_b200._lib import ACT_GELU, ACT_QUICK_GELU
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    tok, pos = torch.randn(50, 64, generator=g), torch.randn(16, 64, generator=g)
    ids = torch.randint(0, 50, (3, 16), generator=g)
    out = ops.embed_tokens(ids.to(dev, torch.int32), tok.to(dev), pos.to(dev)).float().cpu()
    ref = (tok[ids] + pos[None]).to(torch.bfloat16).float().view(48, 64)
    assert torch.equal(out, ref)
    x = torch.randn(40, 72, generator=g).to(torch.bfloat16)
    idx = torch.tensor([3, 39, 0, 17], dtype=torch.int32)
    assert torch.equal(ops.gather_rows(x.to(dev), idx.to(dev)).cpu(), x.float()[idx.long()])
    a = torch.randn(154, 128, generator=g).to(torch.bfloat16)
    w = (torch.randn(256, 128, generator=g) / 11.3).to(torch.bfloat16)
    bias = torch.randn(256, generator=g)
    pre = a.float() @ w.float().t() + bias
    for act, fn in ((ACT_GELU, F.gelu), (ACT_QUICK_GELU, C.quick_gelu)):
        y = ops.gemm(a.to(dev), w.to(dev), bias=bias.to(dev), act=act)
        _check(f"gemm_act{act}", y, fn(pre), 1.5 * 2 ** -9)


def _load(mod, sd):
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.endswith("logit_scale") for m in missing), missing


@gpu
@pytest.mark.parametrize("size", ["tiny", "clip_l"])
def test_frozen_clip_embedder_vs_oracle(size):
    from custom_diffusion360_b200.sgm.modules.encoders.modules import FrozenCLIPEmbedder
    dev = torch.device("cuda:0")
    cfg = dict(TINY_CLIP) if size == "tiny" else dict(C.CLIP_L_CFG)
    sd = C.synthetic_state_dict(C.clip_param_shapes(cfg), seed=5)
    emb = FrozenCLIPEmbedder(layer="hidden", layer_idx=11, arch=cfg)
    assert {k: tuple(v.shape) for k, v in emb.state_dict().items()} == C.clip_param_shapes(cfg)
    _load(emb, sd)
    emb = emb.to(dev)
    tokens = _tokens(cfg, 3, seed=2)
    z = emb(tokens)
    with torch.no_grad():
        if size == "tiny":
            ref = C.clip_text_hidden(sd, cfg, tokens)
        else:
            with torch.device(dev):
                ref = C.clip_text_hidden({k: v.to(dev) for k, v in sd.items()}, cfg, tokens.to(dev))
    assert z.shape == ref.shape == (3, cfg["ctx"], cfg["width"])
    _check(f"frozen_clip_{size}", z, ref, 1.5e-2 if size == "tiny" else 2.5e-2)


@gpu
@pytest.mark.parametrize("size", ["tiny", "bigG"])
def test_frozen_open_clip_embedder_vs_oracle(size):
    from custom_diffusion360_b200.sgm.modules.encoders.modules import FrozenOpenCLIPEmbedder
    dev = torch.device("cuda:0")
    cfg = dict(TINY_OC) if size == "tiny" else dict(C.OPEN_CLIP_BIGG_CFG)
    sd = C.synthetic_state_dict(C.open_clip_param_shapes(cfg), seed=3)
    emb = FrozenOpenCLIPEmbedder(arch=cfg, layer="penultimate", always_return_pooled=True, legacy=False)
    shapes = {k: tuple(v.shape) for k, v in emb.state_dict().items() if not k.endswith("logit_scale")}
    assert shapes == C.open_clip_param_shapes(cfg)
    _load(emb, sd)
    emb = emb.to(dev)
    tokens = _tokens(cfg, 2, seed=4)
    pen, pooled = emb(tokens)
    with torch.no_grad():
        if size == "tiny":
            ref = C.open_clip_text(sd, cfg, tokens)
        else:
            torch.backends.cuda.matmul.allow_tf32 = False
            with torch.device(dev):
                ref = C.open_clip_text({k: v.to(dev) for k, v in sd.items()}, cfg, tokens.to(dev))
    tol = 1.5e-2 if size == "tiny" else 3e-2
    _check(f"open_clip_{size}_penultimate", pen, ref["penultimate"], tol)
    _check(f"open_clip_{size}_pooled", pooled, ref["pooled"], tol)
    # legacy / last variants of the same tower
    emb.legacy, emb.return_pooled, emb.layer = True, False, "last"
    with torch.no_grad():
        w = cfg["width"]
        g = lambda k: sd["model." + k].to(ref["last"].device)
        last_ln = F.layer_norm(ref["last"], (w,), g("ln_final.weight"), g("ln_final.bias"), cfg["eps"])
    _check(f"open_clip_{size}_legacy_last", emb(tokens), last_ln, tol)


@gpu
def test_general_conditioner_sdxl_wiring_vs_oracle():
    """The five embedders of train_co3d_concept.yaml:56-96 (tiny towers): crossattn = [CLIP | OpenCLIP
    penultimate], vector = [pooled | 3 x size embeddings], reference halves appended on the batch axis;
    sample.py's unconditional conditioning (all-zero uc, force_ref_zero_embeddings)."""
    from custom_diffusion360_b200.sgm.modules.encoders.modules import GeneralConditioner
    dev = torch.device("cuda:0")
    clip_cfg, oc_cfg = dict(TINY_CLIP), dict(TINY_OC)
    clip_sd = C.synthetic_state_dict(C.clip_param_shapes(clip_cfg), seed=5)
    oc_sd = C.synthetic_state_dict(C.open_clip_param_shapes(oc_cfg), seed=3)
    size_cfg = lambda k: {"is_trainable": False, "input_keys": f"{k},{k}_ref", "target": E + "ConcatTimestepEmbedderND",
                          "params": {"outdim": 8}}
    cond = GeneralConditioner([
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": E + "FrozenCLIPEmbedder",
         "params": {"layer": "hidden", "layer_idx": 11, "arch": clip_cfg}},
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": E + "FrozenOpenCLIPEmbedder",
         "params": {"arch": oc_cfg, "layer": "penultimate", "always_return_pooled": True, "legacy": False}},
        size_cfg("original_size_as_tuple"), size_cfg("crop_coords_top_left"), size_cfg("target_size_as_tuple")])
    _load(cond.embedders[0], clip_sd)
    _load(cond.embedders[1], oc_sd)
    cond = cond.to(dev)
    b = 2
    size = lambda v, n: torch.tensor([v]).repeat(n, 1)
    t1, t1r = _tokens(clip_cfg, b, 1), _tokens(clip_cfg, 4 * b, 3)
    t2, t2r = _tokens(oc_cfg, b, 2), _tokens(oc_cfg, 4 * b, 4)
    sizes = {"original_size_as_tuple": size([512.0, 512.0], b), "original_size_as_tuple_ref": size([512.0, 384.0], 4 * b),
             "crop_coords_top_left": size([0.0, 16.0], b), "crop_coords_top_left_ref": size([8.0, 0.0], 4 * b),
             "target_size_as_tuple": size([512.0, 512.0], b), "target_size_as_tuple_ref": size([512.0, 512.0], 4 * b)}
    # both text embedders read `txt` / `txt_ref` (the shipped yaml): the value carries the ids of both tokenisers
    batch = dict(txt=(t1, t2), txt_ref=(t1r, t2r), **{k: v.to(dev) for k, v in sizes.items()})
    emb_o = C.sdxl_conditioner(clip_sd, clip_cfg, oc_sd, oc_cfg, size_dim=8)
    batch_o = dict(txt=(t1, t2), txt_ref=(t1r, t2r), **sizes)
    out = cond(batch)
    with torch.no_grad():
        ref = C.general_conditioner(emb_o, batch_o)
    assert out["crossattn"].shape == ref["crossattn"].shape == (b + 4 * b, clip_cfg["ctx"], 256)
    assert out["vector"].shape == ref["vector"].shape == (b + 4 * b, 48 + 6 * 8)
    _check("general_conditioner_crossattn", out["crossattn"], ref["crossattn"], 1.5e-2)
    _check("general_conditioner_vector", out["vector"], ref["vector"], 1.5e-2)
    keys = [e.input_keys for e in cond.embedders]
    c, uc = cond.get_unconditional_conditioning(batch, force_uc_zero_embeddings=keys, force_ref_zero_embeddings=True)
    keys_o = [e["input_keys"] for e in emb_o]
    c_o, uc_o = C.get_unconditional_conditioning(emb_o, batch_o, force_uc_zero_embeddings=keys_o,
                                                 force_ref_zero_embeddings=True)
    assert c["crossattn"].shape == c_o["crossattn"].shape == (b, clip_cfg["ctx"], 256)
    _check("general_conditioner_c_crossattn", c["crossattn"], c_o["crossattn"], 1.5e-2)
    _check("general_conditioner_c_vector", c["vector"], c_o["vector"], 1.5e-2)
    assert float(uc["crossattn"].abs().max()) == 0.0 and float(uc["vector"].abs().max()) == 0.0
    assert float(uc_o["crossattn"].abs().max()) == 0.0


@gpu
def test_modifier_token_rows():
    """`<new1>` (modules.py:418-431, 676-690; main.py:623-624): one appended row per tower initialised from
    row 42170 (mod vocabulary), the only trainable parameter, exported / re-imported as `embed`."""
    from custom_diffusion360_b200.sgm.modules.encoders.modules import GeneralConditioner
    dev = torch.device("cuda:0")
    cond = GeneralConditioner([
        {"is_trainable": False, "input_keys": "txt,txt_ref", "target": E + "FrozenCLIPEmbedder",
         "params": {"layer": "hidden", "layer_idx": 11, "arch": dict(TINY_CLIP), "modifier_token": "<new1>"}},
        {"is_trainable": False, "input_keys": "txt2,txt2_ref", "target": E + "FrozenOpenCLIPEmbedder",
         "params": {"arch": dict(TINY_OC), "layer": "penultimate", "always_return_pooled": True, "legacy": False,
                    "modifier_token": "<new1>"}}]).to(dev)
    e0, e1 = cond.embedders
    w0 = e0.transformer.text_model.embeddings.token_embedding.weight
    assert w0.shape[0] == TINY_CLIP["vocab"] + 1 and e0.modifier_token_id == [TINY_CLIP["vocab"]]
    assert torch.equal(w0[-1], w0[42170 % TINY_CLIP["vocab"]])
    assert [n for n, p in e0.named_parameters() if p.requires_grad] == ["transformer.text_model.embeddings.token_embedding.weight"]
    assert [n for n, p in e1.named_parameters() if p.requires_grad] == ["model.token_embedding.weight"]
    rows = cond.modifier_token_rows()
    assert rows[0].shape == (1, 128) and rows[1].shape == (1, 128)
    # the new token id is embedded from the appended row
    t = _tokens(TINY_CLIP, 1, 0)
    t[0, 1] = TINY_CLIP["vocab"]
    z1 = e0(t)
    cond.load_modifier_token_rows([rows[0] + 1.0, rows[1]])
    z2 = e0(t)
    assert float((z1 - z2).abs().max()) > 1e-3
