#!/usr/bin/env python
"""Per-key-tile timeline of ONE CTA of the ping-pong attention kernel (level-1 self-attention shape).

Builds attention_tcgen05.cu (+ the GEMM file for the tensor-map encoder) with -DCD360_ATT_TRACE into a
private .so, runs the shape once warm and prints, per key tile j and query-tile group t, SM-clock
stamps relative to the CTA's first stamp:
  S_t(j) issued | PV_t(j) issued | group t: loop top, s_full passed, S read (s_free), exps done,
  o_full passed | K/V tile j load issued
"""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_diffusion360_b200 import _build  # noqa: E402


def build():
    out = os.path.join(_build.BUILD, "libcd360_atttrace.so")
    os.makedirs(_build.BUILD, exist_ok=True)
    flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    srcs = [os.path.join(_build.CSRC, f) for f in ("attention_tcgen05.cu", "gemm_tcgen05.cu")]
    subprocess.run([_build._nvcc(), *flags, "-DCD360_ATT_TRACE", "-shared", *srcs, "-o", out], check=True)
    return out


def main():
    lib = C.CDLL(build())
    if not torch.cuda.is_available():
        print("built")
        return
    dev = torch.device("cuda:0")
    lib.cd360_att_set_trace.argtypes = [C.c_void_p]
    L = C.c_int64
    lib.cd360_attention_bf16.argtypes = [C.c_void_p, L, C.c_void_p, L, C.c_void_p, L, C.c_void_p, L,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    shapes = [(3, 10, 4096, 4096), (3, 20, 1024, 1024)]
    for b, h, nq, nkv in shapes:
        c = h * 64
        qkv = torch.randn(b * nq, 3 * c, device=dev).to(torch.bfloat16)
        o = torch.empty(b * nq, c, device=dev, dtype=torch.bfloat16)
        tr = torch.zeros(16 * 64, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
        for it in range(3):
            if it == 2:
                assert lib.cd360_att_set_trace(tr.data_ptr()) == 0
            rc = lib.cd360_attention_bf16(q.data_ptr(), 3 * c, k.data_ptr(), 3 * c, v.data_ptr(), 3 * c,
                                          o.data_ptr(), c, b, h, nq, nkv, st)
            assert rc == 0
        torch.cuda.synchronize()
        lib.cd360_att_set_trace(None)
        t = tr.view(16, 64).cpu()
        nt = nkv // 128
        t0 = int(t[t > 0].min())
        rel = lambda x: int(x) - t0 if int(x) > 0 else -1
        print(f"== b{b} h{h} {nq}x{nkv}: {nt} key tiles; clocks relative to the CTA's first stamp")
        print("  j |  S0 iss  S1 iss  PV0 iss PV1 iss | g0: top   s_full  s_free  exp_done o_full | g1: top   s_full  s_free  exp_done o_full | kv load")
        for j in range(nt):
            row = [rel(t[s, j]) for s in range(16)]
            print(f"{j:3d} | {row[0]:7d} {row[1]:7d} {row[2]:7d} {row[3]:7d} | {row[8]:7d} {row[4]:7d} {row[5]:7d} {row[6]:8d} {row[7]:6d} |"
                  f" {row[13]:7d} {row[9]:7d} {row[10]:7d} {row[11]:8d} {row[12]:6d} | {row[14]:7d}")


if __name__ == "__main__":
    main()
