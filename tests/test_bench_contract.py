"""bench.py contract checks that need no GPU: the reference arm (the reference source of oracle/_ref, or the CPU oracle port, timed on the host cores) prints
exactly one JSON line with the agreed keys, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--ref-frames", "4", "--distinct", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BEV frames/sec (HDL-64E)" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("BASELINE configs[1]: HDL_64E") and d["sample_frames_per_step"] == 4
    assert d["config"]["frames_per_step_per_gpu"] == 4440 and d["config"]["frames_per_wave"] == 2220      # the product arm's config, word for word
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert cb["port_frames_per_s"] > 0
    if cb["kind"] == "reference":      # oracle/_ref present: the line's value is the reference's own source, one process per core
        rs = cb["reference_source"]
        assert rs["procs"] == cb["cores"] and rs["frames"] > 0 and abs(rs["frames_per_s"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
