#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the benchmark step (no nsys in this image).

Builds every csrc/*.cu a second time with -DCD360_TIMELINE into a private library
(custom_diffusion360_b200/_build/libcd360_timeline.so, never the product .so), loads it in place
of libcd360.so, runs bench.py's workload, and after a warm graph replay reads the ring the kernels
wrote: earliest CTA entry / latest exit per launch (ns, %globaltimer, 32 ns..1 us granularity).

Prints per-kernel-kind busy time, launch counts, the summed idle gaps between consecutive kernels
and the longest kernels; writes the raw records to gpurun_out/step_timeline.json.

    python tools/step_timeline.py [--latent 128] [--n-img 1]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_diffusion360_b200 import _build, _lib  # noqa: E402

KINDS = ["gemm", "attention", "groupnorm_stats", "groupnorm_apply", "layernorm", "small_linear",
         "timestep_embedding", "im2col_nchw", "im2col_s2", "upsample2x", "cast_f32_bf16", "cast_bf16_f32",
         "nhwc_to_nchw", "nchw_to_nhwc", "cfg_euler", "cfg_euler_dev", "nerf_points", "nerf_combine",
         "nerf_volrender"]
RING = 8192
SETTERS = ["gemm", "attention", "norm", "elementwise", "nerf"]


def build_timeline_lib():
    out = os.path.join(_build.BUILD, "libcd360_timeline.so")
    os.makedirs(_build.BUILD, exist_ok=True)
    srcs = [os.path.join(_build.CSRC, f) for f in _build.SOURCES]
    deps = srcs + [os.path.join(_build.CSRC, "cd360_common.cuh")]
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]

    def one(src):
        obj = os.path.join(_build.BUILD, "tl_" + os.path.basename(src).replace(".cu", ".o"))
        subprocess.run([_build._nvcc(), *flags, "-DCD360_TIMELINE", "-c", src, "-o", obj], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(one, srcs))
    subprocess.run([_build._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                   check=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--n-img", dest="n_img", type=int, default=1)
    args = ap.parse_args()
    path = build_timeline_lib()
    if not torch.cuda.is_available():
        print("built", path)
        return
    _lib.LIB_PATH = path  # every ops.* call now goes through the instrumented build
    lib = _lib.load(build_if_missing=False)
    dev = torch.device("cuda:0")
    ring = torch.zeros(RING * 3, dtype=torch.int64, device=dev)  # {start, end, kind|ctas<<32}
    for name in SETTERS:
        fn = getattr(lib, f"cd360_tl_set_{name}")
        fn.argtypes = [C.c_void_p]
        assert fn(ring.data_ptr()) == 0

    sys.path.insert(0, ROOT)
    import bench  # noqa: E402

    engine, net, step, x_init, sigmas = bench.make_step(args.latent, args.n_img, dev, 0, True)
    x = x_init.clone()
    with torch.no_grad():
        step(x, float(sigmas[0]), float(sigmas[1]))
        for i in range(1, 5):
            step(x, float(sigmas[i]), float(sigmas[i + 1]))
        torch.cuda.synchronize()
        r3 = ring.view(RING, 3)
        r3.zero_()
        r3[:, 0] = torch.iinfo(torch.int64).max  # atomicMin target (unsigned compare: 0x7fff.. is large enough)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(x, float(sigmas[5]), float(sigmas[6]))
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1)
    rec = r3.cpu()
    used = rec[:, 1] > 0
    rows = []
    for s, e, kc in rec[used].tolist():
        rows.append({"start": s, "end": e, "kind": KINDS[kc & 0xFFFFFFFF], "ctas": kc >> 32})
    rows.sort(key=lambda r: r["start"])
    t0 = rows[0]["start"]
    for r in rows:
        r["start"] -= t0
        r["end"] -= t0
    span = max(r["end"] for r in rows)
    busy = {}
    gaps = 0
    overlap = 0
    prev_end = 0
    gap_by_next = {}
    for r in rows:
        d = r["end"] - r["start"]
        b = busy.setdefault(r["kind"], [0, 0])
        b[0] += 1
        b[1] += d
        g = r["start"] - prev_end
        if g > 0:
            gaps += g
            gg = gap_by_next.setdefault(r["kind"], [0, 0])
            gg[0] += 1
            gg[1] += g
        else:
            overlap += -g
        prev_end = max(prev_end, r["end"])
    summary = {
        "step_ms_events": round(step_ms, 3), "span_ms": round(span / 1e6, 3), "launches": len(rows),
        "idle_gap_ms": round(gaps / 1e6, 3), "overlap_ms": round(overlap / 1e6, 3),
        "busy_ms_by_kind": {k: {"launches": v[0], "ms": round(v[1] / 1e6, 3), "avg_us": round(v[1] / v[0] / 1e3, 2)}
                            for k, v in sorted(busy.items(), key=lambda kv: -kv[1][1])},
        "gap_before_kind": {k: {"count": v[0], "ms": round(v[1] / 1e6, 3), "avg_us": round(v[1] / v[0] / 1e3, 2)}
                            for k, v in sorted(gap_by_next.items(), key=lambda kv: -kv[1][1])},
    }
    print(json.dumps(summary, indent=1))
    # duration histogram of the GEMM launches (how much of the step sits in short kernels)
    gd = sorted(r["end"] - r["start"] for r in rows if r["kind"] == "gemm")
    bins = [(0, 10), (10, 15), (15, 20), (20, 30), (30, 50), (50, 100), (100, 1e9)]
    hist = []
    for lo, hi in bins:
        sel = [d for d in gd if lo * 1e3 <= d < hi * 1e3]
        hist.append({"us": f"{lo}-{hi if hi < 1e9 else 'inf'}", "launches": len(sel), "ms": round(sum(sel) / 1e6, 3)})
    print(json.dumps({"gemm_duration_histogram": hist}))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/step_timeline.json", "w") as f:
        json.dump({"summary": summary, "gemm_hist": hist, "records": rows}, f)


if __name__ == "__main__":
    main()
