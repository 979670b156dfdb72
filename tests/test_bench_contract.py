"""bench.py's JSON contract on CPU: the `--impl reference` line (keys the driver reads; only rank 0
prints) and the argument defaults; the CUDA arm itself is exercised on the GPU box."""
import argparse
import json

import bench


def _args(**kw):
    base = dict(workload="sample", gpus=1, steps=2, warmup=0, impl="reference", latent=128, n_img=1,
                no_cpu_baseline=False, no_graph=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_reference_arm_line(monkeypatch, capsys):
    fake = {"value": 0.05, "unit": bench.UNIT, "cores": 16, "kind": "port", "sample": "oracle batch-3 guided step"}
    monkeypatch.setattr(bench, "cpu_reference", lambda steps, warmup, latent=128: (dict(fake), 20.0, 1))
    bench.run_reference(_args(), rank=0)
    line = json.loads(capsys.readouterr().out.strip())
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] == 16
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("sample.py car0") and "model" not in line["config"]
    bench.run_reference(_args(gpus=2), rank=1)          # torchrun: the other ranks exit without work
    assert capsys.readouterr().out == ""


def test_ncu_traffic_reads_the_committed_summary():
    t = bench.ncu_traffic("gemm_bf16_tcgen05_kernel")
    assert t is not None and 1e6 < t < 1e9               # DRAM bytes per launch, profiles/launches_r02_summary.json
    assert bench.ncu_traffic("no_such_kernel") is None
