"""bench.py's reference arm runs on the CPU: check the JSON line it prints against the contract (keys, units, arm markers)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("Msamples/sec (rtcamp6 scene, 1920x1080)")
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_ours_arm_refuses_to_run_without_a_device():
    """No CPU fallback on the product path: without a GPU the bench exits with an error instead of printing a number."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "metric" not in out.stdout
