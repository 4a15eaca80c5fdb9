"""CPU tests of the bench.py contract the driver relies on: the reference arm (the CPU port of
the reference's path, timed on the host cores) prints one JSON line with the agreed keys, and
the product arm fails loudly without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "3", "--cpu-sample", "192")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "particle-steps/s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3
    assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric
    assert "4096x4096" in d["config"]["workload"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "192x192" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_product_arm_has_no_cpu_fallback():
    """Without an sm_100 device the b200 arm must stop with the library's error, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm runs")
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert "no CPU fallback" in r.stderr


def test_bench_c5_grid_is_the_tests_grid():
    """bench.py builds the tripolar + land grid of its strong_c5 leg without touching tests/ or oracle/ (host mirror of
    make_boundaries); it must be the grid profiles/bench_configs.py and the GPU config tests build through the oracle"""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import bench
    from bench_configs import tripolar
    for nx, ny in ((432, 384), (270, 240)):
        a, b = bench.c5_grid(nx, ny), tripolar(nx, ny, True)
        assert a["bx"] == b["bx"] and a["by"] == b["by"]
        assert np.array_equal(a["mask"], b["mask"]) and a["mask"].dtype == np.uint8
        assert np.array_equal(a["M"], b["M"]) and np.array_equal(a["pc"], b["pc"])
