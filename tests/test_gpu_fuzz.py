"""Seeded fuzz of the CUDA path against the oracle (tests/gpu_fuzz.py): organised scene frames with random ground planes, walls,
dropouts, duplicates and -1 markers, unstructured frames of random size, hot-cell frames, real synthetic frames - all three
sensors, through bevgen_process_host and through the compact staging format after host expansion.  Bit-exact or fail."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_seeded_fuzz_both_staging_formats():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_fuzz.py"), "12", "500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok: 96 frames" in r.stdout
