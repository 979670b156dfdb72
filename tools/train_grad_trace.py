"""Activation-gradient trace (GPU box): dL/dh after the middle block and after every decoder block,
CUDA backward walk vs autograd of the CPU oracle.  Diagnostic tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sgm_oracle as O, train_oracle as T
from tests import test_train_step_gpu as G

dev = torch.device("cuda:0")
cfg = dict(O.TINY_CFG)
sd = O.synthetic_state_dict(cfg, seed=2)
batch = T.synthetic_train_batch(cfg, 16, n_views=3, b=1, seed=5, image=48)
probe_h = {}
ob = dict(batch, _probe_h=probe_h)
names = T.pose_param_names(sd)
sdg = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
total, _ = T.training_loss(sdg, cfg, ob)
total.backward()
eng = G._engine(cfg, sd, dev)
eng.global_step = 1
trace = []
eng.model.diffusion_model.__dict__["_grad_trace"] = trace
opt = eng.configure_optimizers()
loss = eng.training_step(G._to_engine_batch(batch, dev))
print("loss", float(loss), float(total))
for tag, g in trace:
    ref = probe_h[tag].grad                       # [b, c, h, w]
    b, c, h, w = ref.shape
    ours = g.float().cpu().reshape(b, h, w, c).permute(0, 3, 1, 2)
    rel = float((ours - ref).norm() / ref.norm())
    cos = float((ours * ref).sum() / (ours.norm() * ref.norm()))
    print(f"{tag:8s} c={c:4d} hw={h*w:4d} rel_rms {rel:.4f} cosine {cos:.5f} |ref| {float(ref.norm()):.3g}")
