"""Reduce a raw `ncu --csv` launch list of bench.py to ONE steady-state step of the sampling path:
launch index inside the step, kernel, grid, block, duration (ns) and DRAM bytes.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        -k regex:"gemm_bf16|attention_|groupnorm|small_linear|im2col|upsample2x|cfg_euler|timestep_emb|cast_|nerf_|layernorm|splitk" -c 4500 --csv --log-file gpurun_out/ncu_launches_raw.csv \\
        python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
    python tools/ncu_step_list.py gpurun_out/ncu_launches_raw.csv profiles/launches_r01_step.csv

A step starts at `timestep_embedding_kernel`; the last COMPLETE window without FeatureNeRF kernels is
kept (step 0 of an image runs FeatureNeRF once).

Training step (`--train`): windows are delimited by `adamw_kernel` (one per optimiser step) and the last
complete one is kept, FeatureNeRF kernels included:

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:... -c 12000 --csv \
        --log-file gpurun_out/ncu_train_raw.csv python bench.py --workload train --steps 1 --warmup 1 --no-graph
    python tools/ncu_step_list.py --train gpurun_out/ncu_train_raw.csv profiles/launches_r01_train_step.csv"""
import csv
import io
import json
import re
import sys


def main(src, dst, train=False):
    text = open(src, errors="replace").read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    launches = {}
    for r in rows:
        i = int(r["ID"])
        d = launches.setdefault(i, dict(kernel=re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1].split("<")[0],
                                        grid=r["Grid Size"], block=r["Block Size"]))
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name.startswith("gpu__time_duration"):
            d["ns"] = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        else:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d[name.split(".")[0]] = val * scale
    order = [launches[i] for i in sorted(launches)]
    if train:
        ends = [i for i, d in enumerate(order) if d["kernel"].startswith("adamw")]
        windows = [order[a + 1:b + 1] for a, b in zip(ends, ends[1:])]
    else:
        starts = [i for i, d in enumerate(order) if d["kernel"].startswith("timestep_embedding")]
        windows = [order[a:b] for a, b in zip(starts, starts[1:])]
        windows = [w for w in windows if not any(d["kernel"].startswith("nerf_") for d in w)]
    step = windows[-1]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch_in_step", "kernel", "grid", "block", "gpu__time_duration_ns", "dram_read_bytes", "dram_write_bytes"])
        for i, d in enumerate(step):
            w.writerow([i, d["kernel"], d["grid"], d["block"], int(d.get("ns", 0)), int(d.get("dram__bytes_read", 0)),
                        int(d.get("dram__bytes_write", 0))])
    tot = sum(d.get("ns", 0) for d in step)
    by = {}
    for d in step:
        k = by.setdefault(d["kernel"], dict(launches=0, ns=0.0, dram=0.0))
        k["launches"] += 1
        k["ns"] += d.get("ns", 0)
        k["dram"] += d.get("dram__bytes_read", 0) + d.get("dram__bytes_write", 0)
    summary = {k: dict(launches=v["launches"], ms=v["ns"] / 1e6, share=v["ns"] / tot, dram_bytes_per_launch=v["dram"] / v["launches"])
               for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ns"])}
    print(json.dumps(dict(launches=len(step), serialized_ms=tot / 1e6, kernels=summary), indent=1))


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if a != "--train"]
    main(argv[0], argv[1], train="--train" in sys.argv)
