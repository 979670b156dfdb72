#!/usr/bin/env python
"""GEMM / conv launches of ONE eager training step grouped by shape: count, CUDA-event time (launches are
serialised: run with CD360_PDL=0 for exclusive times), TFLOP/s.  Which shapes the step's GEMM time sits in.

    CD360_PDL=0 python tools/train_gemm_shapes.py > gpurun_out/train_gemm_shapes.txt
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
from custom_diffusion360_b200 import ops

dev = torch.device("cuda:0")
engine, net, batch = B.make_train(64, 4, dev, 0)
engine.global_step = 1
opt = engine.configure_optimizers()
for _ in range(2):
    opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
torch.cuda.synchronize()

rec = []
_gemm, _conv = ops.gemm, ops.conv3x3


def timed(key, flops, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    rec.append((key, flops, a, b))
    return r


def gemm(a, w, **kw):
    M = kw.get("M") or a.shape[0]
    K = (kw.get("k0") or a.shape[-1]) + ((kw.get("k1") or kw["a1"].shape[-1]) if kw.get("a1") is not None else 0)
    N = w.shape[0]
    tag = "".join(t for t, on in (("+ln", kw.get("ln_stats") is not None), ("+geglu", kw.get("geglu")),
                                  ("+stats", kw.get("stats_out") is not None), ("+f32", kw.get("out_fp32")),
                                  ("+res", kw.get("residual") is not None)) if on)
    return timed(("gemm", M, N, K, tag), 2.0 * M * N * K, lambda: _gemm(a, w, **kw))


def conv3x3(x, w, Bn, H, W, **kw):
    C = x.shape[-1]
    N = w.shape[0]
    return timed(("conv3x3", Bn * H * W, N, 9 * C, ""), 2.0 * Bn * H * W * N * 9 * C, lambda: _conv(x, w, Bn, H, W, **kw))


ops.gemm, ops.conv3x3 = gemm, conv3x3
import custom_diffusion360_b200.sgm.modules.train_path as TP  # noqa: E402  (modules call ops.<name> at run time)
opt.zero_grad(); engine.training_step(dict(batch)); opt.step()
torch.cuda.synchronize()
ops.gemm, ops.conv3x3 = _gemm, _conv
agg = collections.OrderedDict()
for key, fl, a, b in rec:
    t = a.elapsed_time(b)
    c = agg.setdefault(key, [0, 0.0, 0.0])
    c[0] += 1; c[1] += t; c[2] += fl
tot = sum(v[1] for v in agg.values())
print("launch groups %d, launches %d, summed CUDA-event time %.2f ms (eager: includes host launch gaps)" % (len(agg), len(rec), tot))
for key, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%7.2f ms %5.1f%% %4d x %6.1f us  %7.1f TF/s  %s M=%d N=%d K=%d %s" % (
        t, 100 * t / tot, n, 1e3 * t / n, fl / 1e9 / t, key[0], key[1], key[2], key[3], key[4]))
