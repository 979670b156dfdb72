"""Diagnostic (GPU box): after one optimiser step, who is stale — the captured graph or the eager step?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import sgm_oracle as O
from oracle import train_oracle as T
from tests.test_train_step_gpu import _engine, _to_engine_batch
from custom_diffusion360_b200.sgm.models.diffusion import GraphedTrainStep
from custom_diffusion360_b200.sgm.modules.attention import invalidate_all_packed

dev = torch.device("cuda:0")
cfg = dict(O.TINY_CFG)
sd = O.synthetic_state_dict(cfg, seed=3)
batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=2, seed=9, image=32), dev)
batch.pop("rand")
engine = _engine(cfg, sd, dev)
engine.global_step = 1
engine.learning_rate = 1e-3
opt = engine.configure_optimizers()
unet = engine.model.diffusion_model
gs = GraphedTrainStep(engine, opt, batch)
blocks = list(unet.pose_blocks())
gs(batch, step_optimizer=False)
torch.cuda.synchronize()
# graph-pool pack tensors (kept alive by these references)
held = {n: (m.pose_emb_layers.__dict__["_pk"]["w"], dict(m.pose_featurenerf.model._packed)) for n, m in blocks}
opt.step()
gs(batch, step_optimizer=False)
torch.cuda.synchronize()
lg = float(gs.loss)
for n, m in blocks:
    w_graph, nerf_graph = held[n]
    w_now = m.pose_emb_layers.weight.detach().to(torch.bfloat16)
    fresh = m.pose_featurenerf.model.packed()
    diffs = {k: float((nerf_graph[k].float() - fresh[k].float()).abs().max()) for k in ("wg", "w1p", "w2", "wd")}
    print(n, "pose_emb graph-pack vs weights: %.3e" % float((w_graph.float() - w_now.float()).abs().max()), diffs)
invalidate_all_packed(unet, only_trainable=True)
rand = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in gs.rand.items()}
le = float(engine.training_step(dict(batch, rand=rand)))
te = dict(engine.last_loss_dict)
le_all = float(engine.training_step(dict(batch, rand=rand)))
gs.graph.replay()
torch.cuda.synchronize()
lg_again = float(gs.loss)
tg_old = {k: round(float(v), 6) for k, v in gs.terms.items()}
gs2 = GraphedTrainStep(engine, opt, batch)
for k, v in gs.rand.items():
    if torch.is_tensor(v):
        gs2.rand[k].copy_(v)
for a, b_ in zip(gs2.stratified or [], gs.stratified or []):
    for x, y in zip(a["bins"], b_["bins"]):
        x.copy_(y)
gs2.graph.replay()
torch.cuda.synchronize()
lg_new = float(gs2.loss)
print("old graph %.6f (again %.6f) | eager %.6f | eager again %.6f | fresh capture %.6f" % (
    lg, lg_again, le, le_all, lg_new))
print("terms old graph", tg_old)
print("terms eager    ", {k: round(float(v), 6) for k, v in te.items()})
print("terms new graph", {k: round(float(v), 6) for k, v in gs2.terms.items()})
