"""Training step on the CUDA path (DiffusionEngine.training_step -> loss -> explicit backward ->
fused AdamW) against the CPU training oracle (oracle/train_oracle.py, pinned against the
reference's own training code in tests/test_oracle_vs_reference.py) on the same weights, batch and
injected random draws.

Tolerance.  The reference trains in fp32 (autocast dtype float32 in training, openaimodel.py:992);
this path keeps bf16 activations / activation gradients (unit roundoff 2^-9) with fp32
accumulation, fp32 loss / weight gradients / optimiser.  The forward output already differs by
rel_rms ~1.5e-2 from fp32 (tests/test_unet_gpu.py), so the loss gradient is SEEDED with that error
and it grows along the backward walk: measured with tools/train_grad_trace.py, dL/dh is 1.4 % off
right after the output convolution and 5.4 % off at the middle block (profiles/README_r01.md).
Weight gradients inherit the error of the activation gradient they contract (no cancellation
amplification, checked on the oracle), so per trainable tensor we require
    err = ||g - g_oracle|| <= 0.12 ||g_oracle||  and cosine >= 0.99,
    or, for tensors that carry < 10 % of their block's gradient, err <= 0.005 ||g_oracle(block)||
and over ALL trainable values cosine(g, g_oracle) >= 0.998 (measured 0.9993).  The second clause is used
by exactly one tensor in one case (`output_blocks.1.1...nviews.weight`, no jitter, b = 1: rel_rms 0.29 at
0.34 % of its block's gradient norm; <= 0.06 in the jittered cases): the view-softmax logit gradients sum to
zero over the views, dlogit_v = a_v (da_v - sum_u a_u da_u) with da_v = <dS, silu(h_v)>, so when the views
nearly agree the result is a small difference of nearly equal terms and inherits the bf16 rounding of the
STORED forward activations h_v (the reference keeps them in fp32).  Keeping the logit gradients in fp32 all
the way into the weight gradient (tried: fp32-weighted column sum instead of the bf16 GEMM operand) does not
change it — the rounding happens in the forward.  Loss terms agree to
2e-2 relative.  Measured values are written to gpurun_out/train_parity_metrics.json.
"""
import json
import os

import pytest
import torch

from oracle import sgm_oracle as O
from oracle import train_oracle as T

gpu = pytest.mark.gpu
P = "custom_diffusion360_b200.sgm.modules.diffusionmodules."
METRICS = {}


def _engine_config(cfg):
    disc = {"target": P + "discretizer.LegacyDDPMDiscretization"}
    return dict(
        network_config={"target": P + "openaimodel.UNetModel", "params": cfg},
        denoiser_config={"target": P + "denoiser.DiscreteDenoiser", "params": {
            "num_idx": 1000, "weighting_config": {"target": P + "denoiser_weighting.EpsWeighting"},
            "scaling_config": {"target": P + "denoiser_scaling.EpsScaling"}, "discretization_config": disc}},
        loss_fn_config={"target": P + "loss.StandardDiffusionLossImgRef", "params": {
            "sigma_sampler_config": {"target": P + "sigma_sampling.CubicSampling",
                                     "params": {"num_idx": 1000, "discretization_config": disc}},
            "sigma_sampler_config_ref": {"target": P + "sigma_sampling.DiscreteSampling",
                                         "params": {"num_idx": 50, "discretization_config": disc}}}},
        trainkeys="pose", loss_rgb_lambda=5, loss_fg_lambda=10, loss_bg_lambda=10)


def _engine(cfg, sd, dev, **extra):
    from custom_diffusion360_b200.sgm.util import instantiate_from_config
    engine = instantiate_from_config({"target": "custom_diffusion360_b200.sgm.models.diffusion.DiffusionEngine",
                                      "params": {**_engine_config(cfg), **extra}})
    unet = engine.model.diffusion_model
    missing, unexpected = unet.load_state_dict(sd, strict=False)
    assert not unexpected and all("raymarcher" in m for m in missing)
    engine = engine.to(dev)
    engine.denoiser.sigmas = engine.denoiser.sigmas.to(dev)
    return engine


def _to_engine_batch(batch, dev):
    r = batch["rand"]
    rand = {k: v.to(dev) for k, v in r.items() if k != "jitter"}
    rand["sigma_idx"], rand["sigma_ref_idx"] = r["sigma_idx"], r["sigma_ref_idx"]   # index the host tables
    if "jitter" in r:
        rand["jitter"] = r["jitter"]
    out = {"jpg": batch["x"].to(dev), "jpg_ref": batch["x_ref"].to(dev), "pose": batch["cams"].to(dev),
           "mask": batch["mask"].to(dev), "depth": batch["opacity"].to(dev), "rgb": batch["rgb"].to(dev),
           "drop_im": batch["drop_im"].to(dev),
           "cond": {"crossattn": batch["crossattn"].to(dev), "vector": batch["vector"].to(dev)}, "rand": rand}
    if batch.get("mask_ref") is not None:      # data_co3d.py:485: every reference training batch carries it
        out["mask_ref"] = batch["mask_ref"].to(dev)
    return out


def _record(name, **kw):
    METRICS[name] = kw
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "train_parity_metrics.json"), "w") as f:
        json.dump(METRICS, f, indent=1)


@gpu
@pytest.mark.parametrize("jitter,b,mask_ref", [(False, 1, False), (True, 2, False), (True, 2, True)])
def test_training_step_gradients_vs_oracle(jitter, b, mask_ref):
    dev = torch.device("cuda:0")
    cfg = dict(O.TINY_CFG)
    L, n = 16, 3
    sd = O.synthetic_state_dict(cfg, seed=2)
    batch = T.synthetic_train_batch(cfg, L, n_views=n, b=b, seed=5, image=48, jitter=jitter, mask_ref=mask_ref)
    if b > 1:
        batch["drop_im"] = torch.tensor([1.0, 0.0])[:b]     # second sample: reference images dropped
    total_ref, terms_ref, grads_ref = T.training_gradients(sd, cfg, _oracle_batch(batch), cond_grads=True)
    cond_ref = {k[len("cond."):]: grads_ref.pop(k) for k in ("cond.crossattn", "cond.vector")}
    engine = _engine(cfg, sd, dev)
    engine.global_step = 1
    opt = engine.configure_optimizers()
    opt.zero_grad()
    loss = engine.training_step(_to_engine_batch(batch, dev))
    torch.cuda.synchronize()
    terms = engine.last_loss_dict
    assert abs(float(loss) - float(total_ref)) <= 2e-2 * max(1.0, abs(float(total_ref))), (float(loss), float(total_ref))
    for k in ("loss", "loss_fg", "loss_bg", "loss_rgb"):
        assert abs(terms[k] - float(terms_ref[k])) <= 2e-2 * max(0.05, abs(float(terms_ref[k]))), (k, terms[k], float(terms_ref[k]))
    named = dict(engine.model.diffusion_model.named_parameters())
    block_of = lambda k: k.split(".pose")[0]
    block_norm = {}
    for k, g_ref in grads_ref.items():
        block_norm[block_of(k)] = block_norm.get(block_of(k), 0.0) + float((g_ref * g_ref).sum())
    worst = (0.0, None)
    dot = n1 = n2 = 0.0
    for k, g_ref in grads_ref.items():
        g = named[k].grad.detach().float().cpu()
        if "nviews.bias" in k:              # sum_v dlogit_v == 0 analytically: roundoff in the oracle, exact 0 here
            assert float(g.abs().max()) <= 1e-6 and float(g_ref.abs().max()) <= 1e-6, k
            continue
        dot += float((g * g_ref).sum()); n1 += float((g * g).sum()); n2 += float((g_ref * g_ref).sum())
        err = float((g - g_ref).norm())
        rel = err / float(g_ref.norm())
        cos = float((g * g_ref).sum() / (g.norm() * g_ref.norm()).clamp_min(1e-30))
        bn = block_norm[block_of(k)] ** 0.5
        _record(f"jitter={jitter}/b={b}/mask_ref={mask_ref}/{k}", rel_rms=rel, cosine=cos, ref_norm=float(g_ref.norm()), err_over_block=err / bn)
        minor = float(g_ref.norm()) < 0.1 * bn
        assert (rel <= 0.12 and cos >= 0.99) or (minor and err <= 0.005 * bn), \
            f"{k}: rel_rms {rel:.4g}, cosine {cos:.5f}, err / block norm {err / bn:.4g}"
        worst = max(worst, (rel, k))
    gcos = dot / (n1 * n2) ** 0.5
    _record(f"jitter={jitter}/b={b}/mask_ref={mask_ref}/summary", loss=float(loss), loss_oracle=float(total_ref), worst_rel_rms=worst[0],
            worst_tensor=worst[1], tensors=len(grads_ref), global_cosine=gcos)
    assert gcos >= 0.998, gcos
    # conditioning gradients (what the reference's autograd hands to the text encoders for the `<new1>`
    # token rows, diffusion.py:343-356): same bf16-activation-gradient tolerance as the weights; the rows
    # of the reference views are exactly zero (no_grad stream)
    for k, g_ref in cond_ref.items():
        g = engine.last_cond_grads[k].float().cpu()
        assert g.shape == g_ref.shape, (k, g.shape, g_ref.shape)
        assert float(g[b:].abs().max()) == 0.0 and float(g_ref[b:].abs().max()) == 0.0, k
        rel = float((g - g_ref).norm() / g_ref.norm())
        cos = float((g * g_ref).sum() / (g.norm() * g_ref.norm()).clamp_min(1e-30))
        _record(f"jitter={jitter}/b={b}/mask_ref={mask_ref}/cond.{k}", rel_rms=rel, cosine=cos, ref_norm=float(g_ref.norm()))
        assert rel <= 0.12 and cos >= 0.99, f"cond.{k}: rel_rms {rel:.4g}, cosine {cos:.5f}"
    # optimiser: one fused AdamW step over the flat buffer == torch.optim.AdamW semantics on the same gradients
    before = {k: named[k].detach().clone() for k in grads_ref}
    grads = {k: named[k].grad.detach().clone() for k in grads_ref}
    opt.step()
    torch.cuda.synchronize()
    assert engine.global_step == 2
    for k in grads_ref:
        p_ref, _, _ = T.adamw_step(before[k], grads[k], torch.zeros_like(before[k]), torch.zeros_like(before[k]), 1,
                                   lr=engine.learning_rate)
        assert float((named[k].detach() - p_ref).abs().max()) <= 1e-6 + 1e-5 * float(p_ref.abs().max()), k


def _oracle_batch(batch):
    return dict(batch)


@gpu
def test_conditioning_gradients_reach_an_autograd_conditioner():
    """The drop-in case of INTEGRATION.md: `cond` comes from a torch-autograd conditioner (the
    reference's GeneralConditioner with trainable token rows).  training_step must leave, in the
    conditioner's parameter .grad, the contraction of the step's conditioning gradients with the
    conditioner's own Jacobian — i.e. what `loss.backward()` does in the reference."""
    dev = torch.device("cuda:0")
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=2)
    batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=1, seed=5, image=48), dev)
    ca0, vec0 = batch["cond"]["crossattn"], batch["cond"]["vector"]
    row = torch.nn.Parameter(torch.randn(ca0.shape[-1], device=dev) * 0.1)      # a "token embedding row"
    gain = torch.nn.Parameter(torch.ones(vec0.shape[-1], device=dev))
    onehot = torch.zeros(ca0.shape[0], ca0.shape[1], 1, device=dev)
    onehot[:, 5] = 1.0                                                           # the token sits at position 5
    batch["cond"] = {"crossattn": ca0 + onehot * row, "vector": vec0 * gain}
    engine = _engine(cfg, sd, dev)
    engine.global_step = 1
    opt = engine.configure_optimizers()
    opt.zero_grad()
    engine.training_step(batch)
    torch.cuda.synchronize()
    cg = engine.last_cond_grads
    assert row.grad is not None and gain.grad is not None
    exp_row = (cg["crossattn"] * onehot).sum((0, 1))
    exp_gain = (cg["vector"] * vec0).sum(0)
    assert float(row.grad.abs().max()) > 0
    assert torch.allclose(row.grad, exp_row, rtol=1e-5, atol=1e-8) and torch.allclose(gain.grad, exp_gain, rtol=1e-5, atol=1e-8)


@gpu
def test_training_steps_reduce_the_loss():
    """A few optimiser steps on one fixed batch (fixed noise): the total loss must go down —
    end-to-end sign check of every gradient and of the update."""
    dev = torch.device("cuda:0")
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=3)
    batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=1, seed=9, image=32), dev)
    engine = _engine(cfg, sd, dev)
    engine.global_step = 1
    engine.learning_rate = 2e-3
    opt = engine.configure_optimizers()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        losses.append(float(engine.training_step(dict(batch))))
        opt.step()
    assert losses[-1] < losses[0], losses
    _record("loss_curve", losses=losses)


@gpu
def test_graphed_training_step_equals_eager():
    """GraphedTrainStep (one CUDA-graph replay per batch) against the eager step on the same draws:
    same loss, same gradients (up to fp32 atomic ordering in the scatter / column-sum reductions),
    and the replay sees optimiser updates (the trainable packs are rebuilt inside the graph)."""
    from custom_diffusion360_b200.sgm.models.diffusion import GraphedTrainStep
    dev = torch.device("cuda:0")
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=3)
    batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=2, seed=9, image=32), dev)
    batch.pop("rand")
    engine = _engine(cfg, sd, dev)
    engine.global_step = 1
    engine.learning_rate = 1e-3
    opt = engine.configure_optimizers()
    gs = GraphedTrainStep(engine, opt, batch)
    for it in range(2):
        loss_g = gs(batch, step_optimizer=False)
        torch.cuda.synchronize()
        g_graph = opt.flat.grad.clone()
        eager = dict(batch, rand={k: (v.clone() if torch.is_tensor(v) else v) for k, v in gs.rand.items()})
        loss_e = engine.training_step(eager)
        torch.cuda.synchronize()
        g_eager = opt.flat.grad.clone()
        assert abs(float(loss_g) - float(loss_e)) <= 1e-5 * max(1.0, abs(float(loss_e))), (it, float(loss_g), float(loss_e))
        err = float((g_graph - g_eager).abs().max())
        assert err <= 1e-4 * float(g_eager.abs().max()), (it, err, float(g_eager.abs().max()))
        opt.step()                       # second iteration: both paths must see the updated pose weights
    _record("graphed_vs_eager", max_abs_grad_diff=err, loss=float(loss_e))


@gpu
def test_multi_stream_step_equals_serial(monkeypatch):
    """The training step with the reference stream / FeatureNeRF forward / pose-weight gradient branch
    on side CUDA streams (default) against the same step fully serialised on one stream: identical
    loss, gradients equal up to the order of the fp32 atomics of the scatter / column-sum kernels."""
    from custom_diffusion360_b200.sgm.modules.diffusionmodules import openaimodel as U
    dev = torch.device("cuda:0")
    cfg = dict(O.TINY_CFG)
    sd = O.synthetic_state_dict(cfg, seed=3)
    batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=2, seed=9, image=32, jitter=True), dev)
    engine = _engine(cfg, sd, dev)
    engine.global_step = 1
    opt = engine.configure_optimizers()
    unet = engine.model.diffusion_model
    out = {}
    for mode in ("warm-up", "overlap", "serial"):
        monkeypatch.setattr(U, "OVERLAP_REF_STREAM", mode != "serial")
        opt.zero_grad()
        loss = float(engine.training_step(dict(batch)))
        torch.cuda.synchronize()
        out[mode] = (loss, opt.flat.grad.clone())
    assert unet.__dict__.get("_packs_warm") and unet.__dict__.get("_bwd_warm")
    assert unet.__dict__.get("_side_stream") is not None and unet.__dict__.get("_nerf_stream") is not None
    (l1, g1), (l0, g0) = out["overlap"], out["serial"]
    assert l1 == l0 == out["warm-up"][0], (l1, l0, out["warm-up"][0])
    err = float((g1 - g0).abs().max())
    assert err <= 1e-4 * float(g0.abs().max()), (err, float(g0.abs().max()))
    _record("multi_stream_vs_serial", loss=l1, max_abs_grad_diff=err)
