"""The JSON lines bench.py prints are a contract with the driver: the reference arm is run here on the CPU and checked key
by key, and the last product-arm line measured on a B200 (profiles/r01_bench_1gpu_r1n.json, committed) is held to the same
contract, including the arithmetic that ties its numbers together (roofline = algorithmic bytes / kernel time / peak)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


@pytest.fixture(scope="module")
def ref_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--elements", "64", "--cpu-rows", "8"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_line(ref_line):
    d = ref_line
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "laplacian_gs_mask_apply_throughput" and d["unit"] == "GDOF/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None   # BASELINE.md publishes no number
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_committed_product_line_obeys_the_contract(ref_line):
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_1gpu_r1n.json")))
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d) and "impl" not in d
    for k in ("metric", "unit", "higher_is_better", "dtype", "data", "scaling"):
        assert d[k] == ref_line[k], k                       # both arms report the same metric
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["steps"] >= 20
    assert d["config"]["workload"].split(",")[0] == ref_line["config"]["workload"].split(",")[0]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and {"achieved", "peak", "frac", "traffic"} <= set(r)
    ndof = d["config"]["global_dofs"]
    assert abs(r["achieved"] - r["algorithmic_bytes_per_dof"] * ndof / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.5 < r["frac"] < 1.0
    assert abs(r["traffic"] * 1e9 / (r["algorithmic_bytes_per_dof"] * ndof) - 1.0) < 0.02   # ncu DRAM bytes = algorithmic
    assert abs(d["value"] - ndof / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-9 * d["value"]
    assert 0.9 < r["kernel_share_of_step"] <= 1.0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 8 * ndof and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] == 3 * d["steps"]             # strip + seam_x + seam_y per apply
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < d["e2e"]["value"]
