"""Kernel-time table of one training step with torch.profiler (CUPTI): one replay of the captured step
(GraphedTrainStep; `--eager` profiles the eager launches instead).  Device time per kernel name, summed over
the three streams of the step (so the total exceeds the step time where streams overlap)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench as B

dev = torch.device("cuda:0")
engine, net, batch = B.make_train(64, 4, dev, 0)
engine.global_step = 1
opt = engine.configure_optimizers()
for _ in range(2):
    opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
torch.cuda.synchronize()
if "--eager" in sys.argv:
    run = lambda: (opt.zero_grad(), engine.training_step(dict(batch)), opt.step())
else:
    from custom_diffusion360_b200.sgm.models.diffusion import GraphedTrainStep
    gs = GraphedTrainStep(engine, opt, batch)
    for _ in range(2):
        gs(batch)
    run = lambda: gs(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("step (CUDA events, under the profiler) %.2f ms; summed device time %.2f ms" % (e0.elapsed_time(e1), tot / 1e3))
for t, c, k in rows[:45]:
    print("%9.2f ms %6d  %5.1f%%  %6.1f us avg  %s" % (t / 1e3, c, 100 * t / tot, t / max(c, 1), k[:100]))
