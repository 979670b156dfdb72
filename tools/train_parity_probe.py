"""Gradient-parity probe (GPU box): per-tensor and global agreement of the CUDA training step with the
CPU oracle for a few batch variants.  Diagnostic tool, prints a table; not part of the test suite."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgm_oracle as O, train_oracle as T
from tests import test_train_step_gpu as G

dev = torch.device("cuda:0")
cfg = dict(O.TINY_CFG)
sd = O.synthetic_state_dict(cfg, seed=2)
out = {}
for tag, kw, drop in (("b1", dict(b=1, jitter=False), None), ("b1_jit", dict(b=1, jitter=True), None),
                      ("b2", dict(b=2, jitter=False), None), ("b2_drop", dict(b=2, jitter=False), [1.0, 0.0])):
    batch = T.synthetic_train_batch(cfg, 16, n_views=3, seed=5, image=48, **kw)
    if drop is not None:
        batch["drop_im"] = torch.tensor(drop)
    total_ref, terms_ref, grads_ref = T.training_gradients(sd, cfg, dict(batch))
    eng = G._engine(cfg, sd, dev)
    eng.global_step = 1
    opt = eng.configure_optimizers()
    opt.zero_grad()
    loss = eng.training_step(G._to_engine_batch(batch, dev))
    named = dict(eng.model.diffusion_model.named_parameters())
    num = den1 = den2 = 0.0
    rows = []
    for k, gr in grads_ref.items():
        g = named[k].grad.detach().float().cpu()
        num += float((g * gr).sum()); den1 += float((g * g).sum()); den2 += float((gr * gr).sum())
        if "pose_emb" in k:
            c = g.shape[0]
            r = lambda a, b_: float((a - b_).norm() / b_.norm())
            rows.append((k, r(g, gr), r(g[:, :c], gr[:, :c]), r(g[:, c:], gr[:, c:]), float(gr[:, :c].norm()), float(gr[:, c:].norm())))
    print(tag, "loss", float(loss), float(total_ref), "global cosine", num / (den1 * den2) ** 0.5)
    for row in rows:
        print("   %-60s rel %.4f | x-half %.4f r-half %.4f | norms %.3g %.3g" % row)
