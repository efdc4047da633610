"""CPU tests of the host PCD codec (host/pcd_io.h): what pcl::io::loadPCDFile accepts at
BatchMultiBevGen.cpp:730 - ascii, binary, binary_compressed (LZF, field-major), arbitrary field order, "_" padding, foreign
scalar types, POINTS vs WIDTH*HEIGHT - and what savePCDFileBinary writes at :756.  A small C++ probe built from the product
header parses the files; nothing here needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "point-cloud-preprocessing-tools_b200", "host")
NAMES = ("x", "y", "z", "intensity", "row", "col", "t", "label")
DT = dict(x="<f4", y="<f4", z="<f4", intensity="<f4", row="<u2", col="<u2", t="<u4", label="<i2")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("probe") / "pcd_probe")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, os.path.join(ROOT, "tests", "helpers", "pcd_probe.cpp")])
    return exe


def cloud(n, seed=0):
    rng = np.random.default_rng(seed)
    c = dict(x=rng.normal(0, 30, n), y=rng.normal(0, 30, n), z=rng.normal(0, 3, n), intensity=rng.random(n),
             row=rng.integers(0, 64, n), col=rng.integers(0, 2084, n), t=rng.integers(0, 2**32, n), label=rng.integers(-2, 3, n))
    c = {k: np.asarray(v).astype(DT[k]) for k, v in c.items()}
    if n > 4:
        c["x"][0] = np.nan; c["z"][1] = -0.0; c["intensity"][2] = -1.0; c["y"][3] = np.inf
    return c


def run(probe, path, tmp):
    out = os.path.join(tmp, "dump.bin")
    r = subprocess.run([probe, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = dict(kv.split("=") for kv in r.stdout.split())
    n = int(info["n"])
    raw = open(out, "rb").read()
    got, pos = {}, 0
    for k in NAMES:
        a = np.frombuffer(raw, DT[k], n, pos); pos += a.nbytes; got[k] = a
    return info, got, out + ".pcd"


def same(got, want, keys=NAMES):
    for k in keys:
        assert np.array_equal(got[k].view(np.uint8), np.asarray(want[k], DT[k]).view(np.uint8)), k


def header(fields, sizes, types, counts, n, data, width=None, height=1, points=True):
    width = n if width is None else width
    h = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT %d\nVIEWPOINT 0 0 0 1 0 0 0\n" % (
        " ".join(fields), " ".join(map(str, sizes)), " ".join(types), " ".join(map(str, counts)), width, height)
    if points:
        h += "POINTS %d\n" % n
    return (h + "DATA %s\n" % data).encode()


def lzf_literals(b):
    """A valid LZF stream made of literal runs only (ctrl byte < 32 = run of ctrl + 1 bytes)."""
    out = bytearray()
    for i in range(0, len(b), 32):
        chunk = b[i:i + 32]
        out.append(len(chunk) - 1); out += chunk
    return bytes(out)


def lzf_with_backrefs(b):
    """LZF stream that also uses back references: every 16-byte repeat of the previous 16 bytes becomes (len=16, dist=16)."""
    out = bytearray(); i = 0
    while i < len(b):
        if i >= 16 and b[i:i + 16] == b[i - 16:i] and len(b) - i >= 16:
            out += bytes([(7 << 5) | 0, 16 - 2 - 7, 15]); i += 16        # len field 7 + extra byte: total = 7 + (16-9) + 2 = 16; offset 15 -> distance 16
        else:
            run = b[i:i + 8]; out.append(len(run) - 1); out += run; i += len(run)
    return bytes(out)


def test_canonical_binary_is_packed_and_round_trips(probe, tmp_path):
    c = cloud(1000)
    rec = np.zeros(1000, np.dtype([(k, DT[k]) for k in NAMES]))
    for k in NAMES:
        rec[k] = c[k]
    p = str(tmp_path / "a.pcd")
    open(p, "wb").write(header(NAMES, [4, 4, 4, 4, 2, 2, 4, 2], "FFFFUUUI", [1] * 8, 1000, "binary") + rec.tobytes())
    info, got, rt = run(probe, p, str(tmp_path))
    assert info["packed"] == "1" and info["stride"] == "26" and info["off"] == "0,4,8,12,16,18,20,24"
    same(got, c)
    assert open(rt, "rb").read() == open(p, "rb").read()          # savePCDFileBinary layout: byte-identical round trip


def test_ascii_shuffled_fields_and_missing_field(probe, tmp_path):
    c = cloud(200, 1)
    c["x"][0] = 1.5                                                # no NaN text in this file
    c["y"][3] = 2.5
    fields = ("label", "x", "y", "z", "col", "row", "intensity")   # no `t`: stays 0 like pcl::fromPCLPointCloud2 leaves it
    p = str(tmp_path / "b.pcd")
    with open(p, "wb") as f:
        f.write(header(fields, [2, 4, 4, 4, 2, 2, 4], "IFFFUUF", [1] * 7, 200, "ascii"))
        for i in range(200):
            f.write((" ".join(repr(float(c[k][i])) if DT[k] == "<f4" else str(int(c[k][i])) for k in fields) + "\n").encode())
    info, got, _ = run(probe, p, str(tmp_path))
    assert info["packed"] == "0" and info["n"] == "200"
    same(got, c, [k for k in NAMES if k != "t"])
    assert not got["t"].any()


def test_binary_with_padding_reordered_is_packed(probe, tmp_path):
    c = cloud(300, 2)
    fields = ("label", "x", "_", "y", "z", "col", "row", "intensity", "t")
    dt = np.dtype([("label", "<i2"), ("x", "<f4"), ("pad", "V3"), ("y", "<f4"), ("z", "<f4"), ("col", "<u2"), ("row", "<u2"), ("intensity", "<f4"), ("t", "<u4")])
    rec = np.zeros(300, dt)
    for k in NAMES:
        rec[k] = c[k]
    p = str(tmp_path / "c.pcd")
    open(p, "wb").write(header(fields, [2, 4, 1, 4, 4, 2, 2, 4, 4], "IFUFFUUFU", [1, 1, 3, 1, 1, 1, 1, 1, 1], 300, "binary") + rec.tobytes())
    info, got, _ = run(probe, p, str(tmp_path))
    assert info["packed"] == "1" and info["stride"] == "29" and info["off"] == "2,9,13,21,19,17,25,0"
    same(got, c)


def test_foreign_scalar_types_fall_back_to_the_host_parser(probe, tmp_path):
    n = 50
    c = cloud(n, 3)
    dt = np.dtype([("x", "<f8"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("row", "<u1"), ("col", "<u4"), ("label", "<i4")])
    rec = np.zeros(n, dt)
    c["x"][0] = 3.25; c["row"] = (c["row"] % 200).astype("<u2")
    for k in dt.names:
        rec[k] = c[k]
    p = str(tmp_path / "d.pcd")
    open(p, "wb").write(header(dt.names, [8, 4, 4, 4, 1, 4, 4], "FFFFUUI", [1] * 7, n, "binary") + rec.tobytes())
    info, got, _ = run(probe, p, str(tmp_path))
    assert info["packed"] == "0"                                  # x is f64, row u8, col u32, label i32: values are converted on the host
    same(got, c, [k for k in NAMES if k != "t"])


@pytest.mark.parametrize("compress", [lzf_literals, lzf_with_backrefs])
def test_binary_compressed_field_major(probe, tmp_path, compress):
    n = 400
    c = cloud(n, 4)
    c["label"][:] = -2; c["intensity"][100:300] = 0.5            # repeats, so the back-reference encoder has something to do
    body = b"".join(np.asarray(c[k], DT[k]).tobytes() for k in NAMES)     # field-major (SoA) payload
    comp = compress(body)
    p = str(tmp_path / "e.pcd")
    open(p, "wb").write(header(NAMES, [4, 4, 4, 4, 2, 2, 4, 2], "FFFFUUUI", [1] * 8, n, "binary_compressed") +
                        np.array([len(comp), len(body)], "<u4").tobytes() + comp)
    info, got, _ = run(probe, p, str(tmp_path))
    assert info["packed"] == "0" and info["n"] == str(n)
    same(got, c)


def test_points_wins_over_width_height_and_short_payload_truncates(probe, tmp_path):
    c = cloud(64, 5)
    rec = np.zeros(64, np.dtype([(k, DT[k]) for k in NAMES]))
    for k in NAMES:
        rec[k] = c[k]
    p = str(tmp_path / "f.pcd")                                    # WIDTH/HEIGHT 0 as KittiPointCloudSelect.cpp:207,455 leaves them
    open(p, "wb").write(header(NAMES, [4, 4, 4, 4, 2, 2, 4, 2], "FFFFUUUI", [1] * 8, 64, "binary", width=0, height=0) + rec.tobytes())
    info, got, _ = run(probe, p, str(tmp_path))
    assert info["n"] == "64"
    same(got, c)
    p2 = str(tmp_path / "g.pcd")                                   # header promises 64 points, the file holds 40
    open(p2, "wb").write(header(NAMES, [4, 4, 4, 4, 2, 2, 4, 2], "FFFFUUUI", [1] * 8, 64, "binary") + rec[:40].tobytes())
    info, got, _ = run(probe, p2, str(tmp_path))
    assert info["n"] == "40"
    same(got, {k: v[:40] for k, v in c.items()})
