"""bench.py's line format, checked on the one arm that runs without a GPU (--impl reference: the CPU restatement on the host
cores), on a small netlist so that the CPU suite stays short; and that the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], capture_output=True, text=True, cwd=ROOT,
                          env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--gates", "20000", "--depth", "40", "--inputs", "256",
               "--outputs", "64", "--cpu-lanes-per-core", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gate_lane_evals_per_s" and d["unit"] == "gate-lane evals/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         cwd=ROOT, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)
