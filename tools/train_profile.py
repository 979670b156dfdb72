"""Kernel-time table of one training step with torch.profiler (CUPTI), eager launches (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench as B

dev = torch.device("cuda:0")
engine, net, batch = B.make_train(64, 4, dev, 0)
opt = engine.configure_optimizers()
for _ in range(2):
    opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device time ms %.2f" % (tot / 1e3))
for t, c, k in rows[:40]:
    print("%9.2f ms %6d  %5.1f%%  %s" % (t / 1e3, c, 100 * t / tot, k[:110]))
