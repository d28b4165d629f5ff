"""bench.py keeps the driver's JSON contract: one line, the BASELINE metric, roofline / cpu_baseline / e2e / clocks objects.
The reference arm (CPU) runs here at a reduced size; our arm runs on the GPU box at a reduced size (the driver runs the real sizes)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--n", "24", "--rays", "10", "--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "rmcrt_ray_cell_steps_per_s" and d["unit"] == "ray-cell steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "origin cells" in cb["sample"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rmcrt.so")):
        assert cb["kind"] == "reference"            # the reference's own per-ray routines when they were built
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_our_arm_line(gpu_lib):
    d = run_bench("--n", "48", "--rays", "16", "--steps", "3", "--warmup", "3", "--configs-n", "64")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "rmcrt_ray_cell_steps_per_s" and d["scaling"] == "weak" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["steps"] == 3 and d["warmup"] == 3 and d["gpu_launches"] == 15 and d["value"] > 1e9
    assert d["config"]["ray_coherence"] == 0            # the headline is the library default: the reference's sampling
    r = d["roofline"]
    assert r["bound"] == "l2" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    bc = d["by_coherence"]
    assert set(bc) == {"0", "1", "2"} and bc["0"]["is_default"] and bc["0"]["ray_cell_steps_per_s"] == d["value"]
    assert all(v["ray_cell_steps_per_s"] > 1e9 for v in bc.values())
    cf = d["configs"]
    assert {"single_level_64_1gpu_coherence0", "single_level_64_1gpu_coherence2", "c4_dataonion_64_16_coherence0", "c4_dataonion_64_16_coherence2",
            "c5_flux_reflect_64_coherence0", "c5_flux_reflect_64_coherence2"} <= set(cf)
    assert all(v["ray_cell_steps_per_s"] > 1e9 and v["ms_per_solve"] > 0 for v in cf.values())
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] == 3 * 4 * 50 ** 3 and e["d2h_bytes_per_step"] == 2 * 8 * 50 ** 3
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
