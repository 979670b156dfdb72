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


def test_committed_bench_lines_carry_every_contract_key():
    """The lines the CUDA arm printed on the B200 boxes (profiles/bench_r02_n{1,2,8}.json): every key of the
    bench contract is present and self-consistent (value = images * steps / time, e2e measured with host
    buffers, roofline fraction = achieved / peak, sub-records of the other BASELINE configs at every N)."""
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(__file__)), "profiles")
    for n in (1, 2, 8):
        line = json.load(open(os.path.join(root, f"bench_r02_n{n}.json")))
        assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["n_gpus"] == n
        assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
        assert line["dtype"] == "bf16" and line["data"].startswith("synthetic")
        assert line["config"]["workload"].startswith("sample.py car0") and "model" not in line["config"]
        assert line["warmup"] >= 3 and line["steps"] >= 1
        assert abs(line["value"] - n * line["config"]["n_img_per_gpu"] * 1e3 / line["ms_per_step"]) < 1e-6 * line["value"]
        e2e = line["e2e"]
        assert e2e["unit"] == bench.UNIT and 0 < e2e["value"] <= line["value"] * 1.02
        assert e2e["h2d_bytes_per_step"] > 4 * 128 * 128 * 4 and e2e["d2h_bytes_per_step"] == 4 * 128 * 128 * 4
        assert line["gpu_launches"] == line["launches_per_step"] * line["steps"] * n > 0
        roof = line["roofline"]
        assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s"
        assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and 0.3 < roof["frac"] < 1.0
        clk = line["clocks"]
        assert clk["sm_mhz"] <= clk["sm_max_mhz"] and not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        for sub in ("config3_n_img4", "config5_sweep", "train_step"):
            assert sub in line and "error" not in line[sub], sub
        assert line["train_step"]["ms_per_step"] > 0 and line["config5_sweep"]["per_pose_latency_ms"] > 0
        if n == 1:
            cb = line["cpu_baseline"]
            assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["unit"] == bench.UNIT and cb["value"] > 0
            assert roof["traffic"] is None or roof["traffic"] > 0
