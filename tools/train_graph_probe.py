"""Diagnostic (GPU box): GraphedTrainStep replay against the eager step on the same draws — prints the
loss terms of both, replay-to-replay and eager-to-eager repeatability and the gradient difference, for
two iterations with an optimiser step in between.  `CD360_PDL=0 python tools/train_graph_probe.py`
repeats it without programmatic dependent launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import sgm_oracle as O
from oracle import train_oracle as T
from tests.test_train_step_gpu import _engine, _to_engine_batch
from custom_diffusion360_b200.sgm.models.diffusion import GraphedTrainStep

dev = torch.device("cuda:0")
b = int(os.environ.get("PROBE_B", "2"))
cfg = dict(O.TINY_CFG)
sd = O.synthetic_state_dict(cfg, seed=3)
batch = _to_engine_batch(T.synthetic_train_batch(cfg, 16, n_views=3, b=b, seed=9, image=32), dev)
batch.pop("rand")
engine = _engine(cfg, sd, dev)
engine.global_step = 1
engine.learning_rate = 1e-3
opt = engine.configure_optimizers()
gs = GraphedTrainStep(engine, opt, batch)


def fl(d):
    return {k: round(float(v), 6) for k, v in d.items()}


for it in range(2):
    gs(batch, step_optimizer=False)
    torch.cuda.synchronize()
    lg, tg = float(gs.loss), fl(gs.terms)
    g_graph = opt.flat.grad.clone()
    gs.graph.replay()
    torch.cuda.synchronize()
    lg2 = float(gs.loss)
    g_graph2 = opt.flat.grad.clone()
    eager = dict(batch, rand={k: (v.clone() if torch.is_tensor(v) else v) for k, v in gs.rand.items()})
    le = float(engine.training_step(dict(eager)))
    torch.cuda.synchronize()
    te = dict(engine.last_loss_dict)
    g_eager = opt.flat.grad.clone()
    le2 = float(engine.training_step(dict(eager)))
    torch.cuda.synchronize()
    g_eager2 = opt.flat.grad.clone()
    gs.graph.replay()
    torch.cuda.synchronize()
    lg3 = float(gs.loss)
    print(f"it {it}: graph {lg:.6f} replay-again {lg2:.6f} after-eager {lg3:.6f} | eager {le:.6f} again {le2:.6f}")
    print("   graph terms", tg)
    print("   eager terms", fl(te))
    gm = float(g_eager.abs().max())
    print("   grad: |graph-eager| %.3e  |graph-graph2| %.3e  |eager-eager2| %.3e  max|g| %.3e" % (
        float((g_graph - g_eager).abs().max()), float((g_graph - g_graph2).abs().max()),
        float((g_eager - g_eager2).abs().max()), gm))
    opt.step()
