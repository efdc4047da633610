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


def test_run_length_deflate_round_trips(tmp_path):
    """imgio::deflate_rle (the PNG layers' encoder): any inflater must give the input back.  Run lengths around every
    boundary of the length codes (3..10, 11/13/.., 227, 257, 258) and of the 258-byte match limit (259, 260, 261, 516, 517),
    mixed literals, an empty input, an all-zero layer, noise."""
    import zlib
    exe = str(tmp_path / "deflate_probe")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "helpers", "deflate_probe.cpp"), "-lz"])
    rng = np.random.default_rng(9)
    cases = [b"", b"\x00", b"ab", bytes(224 * 225), rng.integers(0, 256, 5000).astype(np.uint8).tobytes(), bytes([255]) * 70000]
    runs = bytearray()
    for n in list(range(1, 40)) + [66, 67, 130, 131, 226, 227, 228, 256, 257, 258, 259, 260, 261, 262, 515, 516, 517, 518, 774, 1000]:
        runs += bytes([n % 251]) * n + bytes([(n * 7 + 1) % 256])           # a run, then a different byte
    cases.append(bytes(runs))
    sparse = np.zeros(224 * 225, np.uint8); sparse[rng.integers(0, sparse.size, 900)] = 255
    cases.append(sparse.tobytes())
    for i, data in enumerate(cases):
        src, dst = tmp_path / ("in%d" % i), tmp_path / ("out%d" % i)
        src.write_bytes(data)
        subprocess.check_call([exe, str(src), str(dst)])
        z = dst.read_bytes()
        assert zlib.decompress(z) == data, "case %d (%d bytes)" % (i, len(data))
        if len(data) >= 50000 and data.count(0) + data.count(255) == len(data) and len(set(data)) == 1:
            assert len(z) < len(data) // 100          # a constant layer costs 13 bits per 258 bytes
