#!/usr/bin/env python
"""A handful of launches of the step's dominant kernels at the step's shapes, for ONE `ncu --set full`
capture (profiles/ncu_*_r02.csv):

    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attention_' \
        -o gpurun_out/prof_r02 python tools/ncu_targets.py

`python tools/ncu_targets.py train` launches the training step's own kernels at the step's shapes instead
(profiles/ncu_prof_train_kernels_r02.csv): attention backward (dQ + dK/dV, tcgen05) for level-2 self attention
(b1 h20 256x256), level-1 self attention (b1 h10 1024x1024) and reference_attn (24 576 queries x 77 keys, the
query-split dK/dV variant), the TN-mode weight-gradient GEMM (K = 24 576 sample rows, MN-major operands, split-K),
GroupNorm backward on 8-CTA clusters, LayerNorm backward.

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
    if only == "train":
        r = lambda *shape: torch.randn(*shape, device=dev).to(torch.bfloat16)
        for b, h, nq, nkv in [(1, 20, 256, 256), (1, 10, 1024, 1024), (1, 10, 24576, 77)]:
            c = h * 64
            q, k, v, do = r(b * nq, c), r(b * nkv, c), r(b * nkv, c), r(b * nq, c)
            o = ops.attention(q, k, v, b, h, nq, nkv)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            for _ in range(2):
                ops.attention_bwd(q, k, v, o, do, b, h, nq, nkv, dq=dq, dk=dk, dv=dv)
            torch.cuda.synchronize()
        for K, M, N in [(24576, 640, 208), (1024, 1280, 1280)]:
            a_t, w_t = r(K, M), r(K, N)
            for _ in range(2):
                ops.gemm_tn(a_t, w_t)
            torch.cuda.synchronize()
        for hw, c in [(4096, 320), (256, 1280)]:
            x, dy = r(hw, c), r(hw, c)
            gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
            for _ in range(2):
                ops.groupnorm_bwd(x, gamma, beta, dy, 1, hw, silu=True)
            torch.cuda.synchronize()
        x, dy, gamma = r(1024, 640), r(1024, 640), torch.randn(640, device=dev)
        for _ in range(2):
            ops.layernorm_bwd(x, gamma, dy)
        torch.cuda.synchronize()
        return
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
