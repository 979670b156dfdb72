"""In-situ timing of every attention_bwd call of one SDXL training step, grouped by shape (diagnostic)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from custom_diffusion360_b200 import ops

dev = torch.device("cuda:0")
engine, net, batch = B.make_train(64, 4, dev, 0)
opt = engine.configure_optimizers()
for _ in range(2):
    opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
torch.cuda.synchronize()
recs = []
orig = ops.attention_bwd
def timed(q, k, v, o, dout, batch_, heads, nq, nkv, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = orig(q, k, v, o, dout, batch_, heads, nq, nkv, **kw)
    b.record()
    recs.append(((heads, nq, nkv, kw.get("dk") is not None, k.stride(0)), a, b))
    return r
ops.attention_bwd = timed
import custom_diffusion360_b200.sgm.modules.train_path as TP
opt.zero_grad(); engine.training_step(dict(batch))
torch.cuda.synchronize()
agg = collections.defaultdict(list)
for key, a, b in recs:
    agg[key].append(a.elapsed_time(b) * 1e3)
for key, v in sorted(agg.items()):
    print(key, "calls", len(v), "mean us %.1f" % (sum(v) / len(v)), "min %.1f max %.1f" % (min(v), max(v)), "total ms %.2f" % (sum(v) / 1e3))
