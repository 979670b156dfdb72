#!/usr/bin/env python
"""Phase timeline of the tcgen05 GEMM kernel at the step's shapes.

Builds gemm_tcgen05.cu a second time with -DCD360_GEMM_TRACE (a private .so under
custom_diffusion360_b200/_build/, never the product library), runs each shape with the L2 flushed
and prints, as medians over CTAs in ns relative to the first CTA's entry:

  entry | setup done | first TMA issued | first operands landed | 2nd k-block landed |
  last TMA issued | last MMA committed | accumulator ready (epilogue starts) |
  last store issued | stores drained | exit

    python tools/gemm_trace.py            # needs a GPU; build alone works without
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_diffusion360_b200 import _build, _lib  # noqa: E402
from custom_diffusion360_b200.sgm.prepack import pack_conv3x3, pack_geglu  # noqa: E402

SLOTS = ["entry", "setup", "pdl_wait", "tma_first", "tma_last", "full0", "full1", "mma_last",
         "acc_first", "acc_last", "store_last", "drained", "exit"]


def build_trace_lib():
    out = os.path.join(_build.BUILD, "libcd360_gemmtrace.so")
    os.makedirs(_build.BUILD, exist_ok=True)
    src = os.path.join(_build.CSRC, "gemm_tcgen05.cu")
    if os.path.exists(out) and os.path.getmtime(out) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_build.CSRC, "cd360_common.cuh"))):
        return out
    flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    cmd = [_build._nvcc(), *flags, "-DCD360_GEMM_TRACE", "-shared", src, "-o", out]
    subprocess.run(cmd, check=True)
    return out


def main():
    lib = C.CDLL(build_trace_lib())
    if not torch.cuda.is_available():
        print("built", lib)
        return
    lib.cd360_gemm_bf16.restype = C.c_int
    lib.cd360_gemm_bf16.argtypes = [C.POINTER(_lib.GemmArgs), C.c_void_p]
    lib.cd360_gemm_set_trace.argtypes = [C.c_void_p]
    dev = torch.device("cuda:0")
    trace = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
    assert lib.cd360_gemm_set_trace(trace.data_ptr()) == 0
    flush = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    shapes = [  # M, N, K, geglu, residual, conv(B,H,W,C) or None, what
        (3072, 1280, 1280, False, True, None, "L2 out-proj + residual"),
        (3072, 1280, 1280, False, False, None, "L2 q-proj"),
        (3072, 3840, 1280, False, False, None, "L2 QKV"),
        (3072, 10240, 1280, True, False, None, "L2 FF1 geglu"),
        (3072, 1280, 5120, False, True, None, "L2 FF2 + residual"),
        (12288, 640, 640, False, True, None, "L1 out-proj + residual"),
        (12288, 1920, 640, False, False, None, "L1 QKV"),
        (49152, 320, 2880, False, False, (3, 128, 128, 320), "conv 320->320 @128"),
        (8192, 8192, 8192, False, False, None, "8192^3"),
    ]
    for M, N, K, geglu, use_res, conv, what in shapes:
        if conv is None:
            a = torch.randn(M, K, device=dev).to(torch.bfloat16)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
        else:
            B, H, W, Cc = conv
            a = torch.randn(B, H, W, Cc, device=dev).to(torch.bfloat16)
            w = pack_conv3x3((torch.randn(N, Cc, 3, 3, device=dev) / math.sqrt(K))).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        if geglu:
            w, bias = pack_geglu(w, bias)
        n_out = N // 2 if geglu else N
        res = torch.randn(M, n_out, device=dev).to(torch.bfloat16) if use_res else None
        o = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        g = _lib.GemmArgs()
        g.a0, g.lda0, g.k0 = a.data_ptr(), (K if conv is None else 0), (K if conv is None else 0)
        g.w, g.bias = w.data_ptr(), bias.data_ptr()
        if res is not None:
            g.residual, g.ldr = res.data_ptr(), n_out
        g.out, g.ldo, g.M, g.N = o.data_ptr(), n_out, M, N
        g.geglu = 1 if geglu else 0
        if conv is not None:
            g.conv, g.B, g.H, g.W, g.C = 1, *conv
        for warm_l2 in (False, True):
            rows, crows, evs = [], [], []
            for rep in range(7):
                if warm_l2:
                    lib.cd360_gemm_bf16(C.byref(g), stream)
                else:
                    flush.sum()  # evict L2 with clean lines
                trace.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.cd360_gemm_bf16(C.byref(g), stream)
                e1.record()
                torch.cuda.synchronize()
                assert rc == 0, rc
                evs.append(e0.elapsed_time(e1) * 1e3)
                t = trace.view(148, 32).cpu()
                used = t[:, 0] > 0
                t = t[used].double()
                t0 = t[:, 0].min()
                rel = t - t0
                rel[t == 0] = float("nan")
                rows.append(torch.nanmedian(rel[:, :13], dim=0).values)
                clk = t[:, 16:28] - t[:, 16:17]
                clk[t[:, 16:28] == 0] = float("nan")
                crows.append(torch.nanmedian(clk, dim=0).values)
            med = torch.stack(rows).nanmedian(dim=0).values
            evs.sort()
            rec = {"shape": what, "MNK": [M, N, K], "l2": "warm" if warm_l2 else "flushed",
                   "ctas": int(used.sum()), "event_us": round(evs[len(evs) // 2], 1),
                   "ns": {k: (None if math.isnan(v) else int(v)) for k, v in zip(SLOTS, med.tolist())},
                   # SM clocks since the accumulator of the LAST tile was ready (warp 4): per slab
                   # [math done, res added, store drained, barrier1, store issued], then exit
                   "epi_clk": [None if math.isnan(v) else int(v)
                               for v in torch.stack(crows).nanmedian(dim=0).values.tolist()]}
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
