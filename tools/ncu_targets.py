#!/usr/bin/env python
"""A handful of launches of the step's dominant kernels at the step's shapes, for ONE `ncu --set full`
capture (profiles/ncu_*_r02.csv):

    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attention_' \
        -o gpurun_out/prof_r02 python tools/ncu_targets.py

Order (each 2 launches; the second is L2-warm like inside the step):
  gemm 3072x1280x1280 (+bias +residual, level-2 out projection), gemm 3072x3840x1280 (QKV),
  gemm 3072x1280x5120 (FF2), attention b3 h20 1024x1024 (level-2 self), b3 h10 4096x4096 (level-1 self),
  b3 h20 1024x77 (level-2 text cross).
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_diffusion360_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    torch.manual_seed(0)
    if only in ("", "gemm"):
        for M, N, K in [(3072, 1280, 1280), (3072, 3840, 1280), (3072, 1280, 5120)]:
            a = torch.randn(M, K, device=dev).to(torch.bfloat16)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
            bias = torch.randn(N, device=dev)
            res = torch.randn(M, N, device=dev).to(torch.bfloat16)
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for _ in range(2):
                ops.gemm(a, w, bias=bias, residual=res, out=o)
            torch.cuda.synchronize()
    if only in ("", "attn"):
        for b, h, nq, nkv in [(3, 20, 1024, 1024), (3, 10, 4096, 4096), (3, 20, 1024, 77)]:
            c = h * 64
            q = torch.randn(b * nq, c, device=dev).to(torch.bfloat16)
            k = torch.randn(b * nkv, c, device=dev).to(torch.bfloat16)
            v = torch.randn(b * nkv, c, device=dev).to(torch.bfloat16)
            o = torch.empty(b * nq, c, device=dev, dtype=torch.bfloat16)
            for _ in range(2):
                ops.attention(q, k, v, b, h, nq, nkv, out=o)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
