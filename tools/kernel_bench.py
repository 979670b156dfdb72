#!/usr/bin/env python
"""Per-shape timing of the tensor-core kernels at the shapes of the SDXL step (UNet batch 3,
128x128 latents), CUDA events on the launching stream, L2 flushed between repetitions.

    python tools/kernel_bench.py [--only gemm|conv|attn] [--reps 20]

Prints one JSON line per shape: algorithmic TFLOP/s and fraction of the measured burst peak.
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_diffusion360_b200 import ops  # noqa: E402
from custom_diffusion360_b200.sgm.prepack import pack_conv3x3, pack_geglu  # noqa: E402

dev = torch.device("cuda:0")
flush = None


def timeit(fn, reps, prep=None):
    global flush
    if flush is None:
        flush = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.sum()  # evict L2 by READING 256 MB (a memset would leave it full of dirty lines)
        if prep is not None:
            prep()  # e.g. the producer's write: leaves the input L2-resident as inside the step
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3  # median, min (us)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=15)
    args = ap.parse_args()
    peak = 1693.1
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["bf16_tflops"]
    out = []

    def report(kind, name, flops, us_med, us_min, **kw):
        rec = dict(kind=kind, shape=name, us=round(us_med, 1), us_min=round(us_min, 1),
                   tflops=round(flops / us_med / 1e6, 1), frac=round(flops / us_med / 1e6 / peak, 3), **kw)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    if args.only in ("", "gemm"):
        shapes = [  # (M, N, K, geglu, what)
            (3072, 10240, 1280, True, "L2 FF1 geglu"), (3072, 1280, 5120, False, "L2 FF2"),
            (3072, 3840, 1280, False, "L2 QKV"), (3072, 1280, 1280, False, "L2 out/q proj"),
            (12288, 5120, 640, True, "L1 FF1 geglu"), (12288, 640, 2560, False, "L1 FF2"),
            (12288, 1920, 640, False, "L1 QKV"), (12288, 640, 640, False, "L1 out/q proj"),
            (231, 2560, 2048, False, "L2 ctx KV proj"), (231, 166400, 2048, False, "all-block ctx KV proj"),
            (8192, 8192, 8192, False, "8192^3"),
        ]
        for M, N, K, geglu, what in shapes:
            a = torch.randn(M, K, device=dev).to(torch.bfloat16)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
            bias = torch.randn(N, device=dev)
            res = None if geglu else torch.randn(M, N, device=dev).to(torch.bfloat16)
            if geglu:
                w, bias = pack_geglu(w, bias)
            o = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.bfloat16)
            for cfg in (1024, 512, 128):
                if geglu and cfg == 128:
                    continue
                med, mn = timeit(lambda: ops.gemm(a, w, bias=bias, residual=res, geglu=geglu, out=o, block_n=cfg), args.reps)
                report("gemm", f"{M}x{N}x{K} {what}", 2.0 * M * N * K, med, mn, cfg=cfg)
            del a, w, o, res
    if args.only in ("", "conv"):
        for B, H, C, Co, what in [(3, 128, 320, 320, "L0 320->320"), (3, 128, 960, 320, "L0 dec 960->320"),
                                  (3, 64, 640, 640, "L1 640->640"), (3, 64, 1920, 640, "L1 dec 1920->640"),
                                  (3, 32, 1280, 1280, "L2 1280->1280"), (3, 32, 2560, 1280, "L2 dec 2560->1280")]:
            x = torch.randn(B * H * H, C, device=dev).to(torch.bfloat16)
            w = pack_conv3x3(torch.randn(Co, C, 3, 3, device=dev) / math.sqrt(9 * C))
            bias = torch.randn(Co, device=dev)
            o = torch.empty(B * H * H, Co, device=dev, dtype=torch.bfloat16)
            for cfg in (1024, 512, 128):
                med, mn = timeit(lambda: ops.conv3x3(x, w, B, H, H, bias=bias, out=o, block_n=cfg), args.reps)
                report("conv3x3", f"B{B} {H}x{H} {what}", 2.0 * B * H * H * Co * 9 * C, med, mn, cfg=cfg)
            del x, w, o
    if args.only in ("", "attn"):
        for b, h, nq, nkv, what in [(3, 20, 1024, 1024, "L2 self"), (3, 10, 4096, 4096, "L1 self"),
                                    (3, 20, 1024, 77, "L2 text cross"), (3, 10, 4096, 77, "L1 text cross"),
                                    (3, 10, 98304, 77, "L1 NeRF-sample cross")]:
            c = h * 64
            q = torch.randn(b * nq, c, device=dev).to(torch.bfloat16)
            k = torch.randn(b * nkv, c, device=dev).to(torch.bfloat16)
            v = torch.randn(b * nkv, c, device=dev).to(torch.bfloat16)
            o = torch.empty(b * nq, c, device=dev, dtype=torch.bfloat16)
            med, mn = timeit(lambda: ops.attention(q, k, v, b, h, nq, nkv, out=o), args.reps)
            report("attention", f"b{b} h{h} {nq}x{nkv} {what}", 4.0 * b * h * nq * nkv * 64, med, mn)
            del q, k, v, o
    if args.only in ("", "norm"):
        for B, hw, c0, c1, what in [(3, 16384, 320, 0, "L0 320"), (3, 16384, 640, 320, "L0 dec 640+320"),
                                    (3, 4096, 640, 0, "L1 640"), (3, 4096, 1280, 640, "L1 dec 1280+640"),
                                    (3, 1024, 1280, 0, "L2 1280"), (3, 1024, 1280, 1280, "L2 dec 1280+1280")]:
            x0 = torch.randn(B * hw, c0, device=dev).to(torch.bfloat16)
            x1 = torch.randn(B * hw, c1, device=dev).to(torch.bfloat16) if c1 else None
            src = x0.clone()
            g, bt = torch.randn(c0 + c1, device=dev), torch.randn(c0 + c1, device=dev)
            o = torch.empty(B * hw, c0 + c1, device=dev, dtype=torch.bfloat16)
            nbytes = 2.0 * 2 * B * hw * (c0 + c1)  # read once + write once
            for mode, prep in (("hbm", None), ("l2", lambda: x0.copy_(src))):
                med, mn = timeit(lambda: ops.groupnorm(x0, g, bt, B, hw, x1=x1, out=o), args.reps, prep)
                report("groupnorm+silu", f"b{B} hw{hw} c{c0}+{c1} {what} [{mode}]", 0.0, med, mn,
                       gbps=round(nbytes / med / 1e3, 1))
            del x0, x1, o, src
        for rows, c, what in [(3072, 1280, "L2"), (12288, 640, "L1")]:
            x = torch.randn(rows, c, device=dev).to(torch.bfloat16)
            src = x.clone()
            g, bt = torch.randn(c, device=dev), torch.randn(c, device=dev)
            o = torch.empty_like(x)
            for mode, prep in (("hbm", None), ("l2", lambda: x.copy_(src))):
                med, mn = timeit(lambda: ops.layernorm(x, g, bt, out=o), args.reps, prep)
                report("layernorm", f"{rows}x{c} {what} [{mode}]", 0.0, med, mn,
                       gbps=round(2.0 * 2 * rows * c / med / 1e3, 1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kernel_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
