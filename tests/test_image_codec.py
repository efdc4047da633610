"""CPU tests of the host file encoders (host/image_io.h): the replacements of cv::imwrite / cv::format(FMT_CSV) at
BatchMultiBevGen.cpp:318, :361, :371 and CloudManip.cpp:97-108.  PNG parity = decoded pixels (cv2); CSV parity = text."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "point-cloud-preprocessing-tools_b200", "host")


def test_png_and_csv_encoders(tmp_path):
    cv2 = pytest.importorskip("cv2")
    exe = str(tmp_path / "image_probe")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "helpers", "image_probe.cpp"), "-lz"])
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (224, 224)).astype(np.uint8)
    a[:50] = 0; a[60, :] = 255                                    # sparse like a BEV layer, plus a saturated row
    f = rng.uniform(-3, 300, (201, 201)).astype(np.float32)
    f[0, :6] = [0.5, 1.5, 2.5, -0.5, 254.5, 255.5]                 # cvRound: half to even; saturate at 0 / 255
    f[1, :3] = [np.nan, np.inf, -np.inf]
    f[2, :4] = [0.0, 1e-5, 123456.0, 0.00012345]
    (tmp_path / "a.bin").write_bytes(a.tobytes()); (tmp_path / "f.bin").write_bytes(f.tobytes())
    pre = str(tmp_path / "out")
    subprocess.check_call([exe, str(tmp_path / "a.bin"), str(tmp_path / "f.bin"), pre])
    img = cv2.imread(pre + ".png", cv2.IMREAD_UNCHANGED)
    assert img is not None and img.dtype == np.uint8 and img.shape == (224, 224) and np.array_equal(img, a)
    want = "\n".join(", ".join("%3d" % v for v in row) for row in a) + "\n"           # cv::Formatter::FMT_CSV of CV_8U
    assert open(pre + ".csv").read() == want
    fmt = lambda v: "nan" if np.isnan(v) else ("inf" if v == np.inf else ("-inf" if v == -np.inf else "%.4g" % v))
    want_f = "\n".join(", ".join(fmt(v) for v in row) for row in f) + "\n"              # set32fPrecision(4)
    assert open(pre + "_f.csv").read() == want_f
    png = cv2.imread(pre + "_f.png", cv2.IMREAD_UNCHANGED)
    finite = np.nan_to_num(f, nan=0.0, posinf=1e9, neginf=-1e9)
    u8 = np.clip(np.rint(finite), 0, 255).astype(np.uint8)          # saturate_cast<uchar>(cvRound(v)): half to even, clamp; NaN -> 0
    assert np.array_equal(png, u8)
    m = np.isfinite(f)                                              # OpenCV's own float -> u8 conversion agrees where it is defined
    cvt = cv2.add(np.where(m, f, 0).astype(np.float32), 0, dtype=cv2.CV_8U)
    assert np.array_equal(png[m], cvt[m])
