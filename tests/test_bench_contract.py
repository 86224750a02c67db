"""The driver-facing contract of bench.py: one JSON line with the agreed keys, for the reference arm (CPU, any box)
and for the GPU arm (tiny run, -m gpu)."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"].startswith("env-steps/sec") and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "40", "--warmup", "3", "--no-vecenv", "--no-ppo", "--no-configs", "--sweep", "--rotating-handles", "8",
              "--cpu-budget", "2"])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    # the K-launch block is repeated `reps` times inside one event pair when K launches are shorter than 20 ms
    assert d["n_gpus"] == 1 and d["steps"] == 40 and d["reps"] >= 1 and d["gpu_launches"] == 40 * d["reps"]
    assert d["scaling"] == "weak" and d["dtype"] == "f32" and d["timed_region_s"] >= 0.015
    assert d["pybullet_cpu"].startswith("n/a") and d["e2e_alternative"]["value"] > 0
    assert d["value"] > 1e8 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 4096) < 1e-3 * 4096        # value = envs / time per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4096 * 16 and e["d2h_bytes_per_step"] == 4096 * (13 * 4 + 4 + 1 + 4) and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_roofline_rows_find_the_committed_ncu_captures():
    """bench.py's per-variant roofline rows quote the ncu DRAM bytes of the SAME configuration from profiles/ncu_summary_r*.json:
    the lookup keys (envs, substeps, variant prefix) must match what tools/summarize_profiles*.py wrote."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.ncu_traffic(4096, 8) == 560640.0                      # headline kernel, 4096 envs
    assert bench.ncu_traffic(4194304, 8) > 1.2e9 and bench.ncu_traffic(4194304, 1) > 1.1e9
    for envs, sub, variant in ((12, 1, "norm_"), (4096, 8, "norm_"), (16384, 8, "phys3_"), (131072, 8, "full_rw3_"),
                               (131072, 8, "full_rw8_"), (131072, 8, "full_rw9_")):
        assert bench.ncu_traffic(envs, sub, variant) is not None, (envs, sub, variant)
    assert bench.ncu_traffic(777, 8) is None and bench.ncu_traffic(4096, 8, "no_such_") is None
    r = bench.roofline(4096, 4e-6, substeps=8)
    assert r["bound"] == "hbm" and r["traffic"] == 560640.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] == 4096 * bench.BYTES_PER_ENV_STEP
