"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU
restatement timed on the host cores) must put exactly one JSON line on stdout with the keys the
driver reads, whatever else gets printed goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-budget", "2"], capture_output=True, text=True, cwd=ROOT, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "bev_pool_samples_per_s" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line_with_roofline():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "16", "--warmup", "3",
                        "--no-cpu-baseline"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["steps"] == 16 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["value"] > 0
    assert d["gpu_launches"] > 0 and d["gpu_launches"] % 16 == 0   # whole steps of the hot path's kernels
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] < d["value"]  # host copies inside the timed region
    assert set(d["stages_ms"]) == {"prepare", "feat_rows", "fwd", "og_rows", "bwd"}
    # the e2e step ships the calibration, not a materialised coor
    assert e2e["h2d_bytes_per_step"] < d["e2e_coor_input"]["h2d_bytes_per_step"] - 40e6
    assert d["clocks"]["samples_pre_spin"] > 20, d["clocks"]
    cfg = d["configs"]
    assert "error" not in cfg, cfg
    for name in ("temporal_8f", "hires_256_B1", "hires_256_B8", "radar_128", "radar_512"):
        assert cfg[name]["samples_per_s"] > 0, (name, cfg[name])
    # the checker-side legs report their own failure instead of taking the line down
    assert "prepare_torch_ms" in d["reference_gpu"] or "error" in d["reference_gpu"], d["reference_gpu"]
    st = d["variants"]["strips"]
    assert st["status"] == 0 and st["fwd_rel_err_vs_cells"] < 1e-5 and st["fwd_ms"] > 0 and st["bwd_ms"] > 0, st
