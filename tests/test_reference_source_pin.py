"""Pins the oracle to the reference's OWN source text.

oracle/_ref/libbevgen_ref.so is /root/reference/BatchMultiBevGen.cpp + src/Utility.cpp compiled unmodified against the
stand-in headers of oracle/stub/ (recipe: oracle/Makefile, target `ref`; see oracle/stub/README.md for what the stubs
restate).  Every expression of getOrderedCloud / markGroundPoints / getBelongingGrid / the two binning loops /
readKeyframePose / selectMajorFrames / getKeyFrameLabel / saveLabels / main() that runs below is the reference's code.
The oracle restatement (oracle/bevgen_oracle.c) must agree with it bit for bit; the GPU tests then compare CUDA with the
oracle on the same inputs (tests/cases.py), and with the golden vectors generated from this library
(tests/golden/make_bev_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import cases
from conftest import FIELDS

SENSORS = ("HDL_32E", "OS1_64", "HDL_64E")


@pytest.fixture(scope="module")
def R(O):
    if O.ref_bevgen_lib() is None:
        pytest.skip("oracle/_ref/libbevgen_ref.so not built (needs /root/reference at build time)")
    return O


def _same_float(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


def check_frame(O, sensor, f, double_libm=False, what=""):
    """oracle == reference source on one frame, every observable: ordered cloud, owner, ground_mat, labels, images, files."""
    sp = O.sensor(sensor)
    args = [f[k] for k in FIELDS]
    n = len(f["x"])
    r = O.ref_frame(sensor, *args, t=np.arange(1, n + 1, dtype=np.uint32), double_libm=double_libm, want_csv=True)
    oc = O.order(sp, *args)
    for k in ("x", "y", "z", "intensity"):
        assert _same_float(r[k], oc[k]), (what, "ordered", k)
    assert np.array_equal(r["t"], oc["owner"]), (what, "owner")          # t = input index + 1: the slot's last writer
    lab, gm1, gmf, avg = O.mark_ground(sp, oc, double_libm=double_libm)
    assert np.array_equal(r["ground_mat"], gmf), (what, "ground_mat", int((r["ground_mat"] != gmf).sum()))
    assert np.array_equal(r["label"], lab), (what, "label", int((r["label"] != lab).sum()))
    o = O.frame(sp, *args, double_libm=double_libm)
    assert np.array_equal(o["label"], lab) and np.array_equal(o["owner"], oc["owner"])
    assert np.array_equal(r["single"], o["single"]), (what, "single")
    assert np.array_equal(r["multi"], o["multi"]), (what, "multi")
    assert r["bin_equal"], (what, ".bin bytes != the layers handed to imwrite")
    want = "\n".join(", ".join("%3d" % v for v in row) for row in o["single"]) + "\n"
    assert r["csv"].decode() == want, (what, "csv")
    return r, o


def test_overload_resolution_of_this_toolchain(R):
    """BatchMultiBevGen.cpp:173,179: with <math.h> in the include tree (VTK <= 8 / OpenCV flann) the unqualified calls
    bind to the float overloads; without it atan2 / sqrt / round bind to the C double functions; abs is float either way
    (BatchMultiBevGen.h:14 includes <stdlib.h>)."""
    assert R.ref_math_overloads(False) == dict(atan2="float", sqrt="float", abs="float", round="float")
    assert R.ref_math_overloads(True) == dict(atan2="double", sqrt="double", abs="float", round="double")


@pytest.mark.parametrize("sensor", SENSORS)
def test_synthetic_frames(R, synth, sensor):
    for idx in (0, 1, 300):
        r, o = check_frame(R, sensor, synth.make_frame(sensor, idx), what="%s/%d" % (sensor, idx))
        assert (r["label"] == 0).sum() > 1000 and r["multi"].any() and r["single"].any()
    if sensor == "HDL_64E":
        check_frame(R, sensor, synth.make_frame(sensor, 100, kitti_quirk=True), what="kitti quirk")


@pytest.mark.parametrize("sensor", SENSORS)
def test_random_unstructured_frames(R, sensor):
    sp = R.sensor(sensor)
    for i, f in enumerate(cases.random_unstructured_frames(sp)):
        check_frame(R, sensor, f, what="%s random %d" % (sensor, i))


@pytest.mark.parametrize("double_libm", [False, True])
def test_borderline_angles_both_overload_sets(R, double_libm):
    """Pairs within +-40 ulp of the 10-degree threshold: the float-overload build must equal the float oracle, the
    double-overload build the -DORACLE_DOUBLE_LIBM oracle, and the two builds must differ somewhere."""
    sp = R.sensor("HDL_32E")
    for i, f in enumerate(cases.borderline_frames(sp)):
        check_frame(R, "HDL_32E", f, double_libm=double_libm, what="borderline %d dbl=%s" % (i, double_libm))
    if double_libm:
        f = cases.borderline_frames(sp)[0]
        a = R.ref_frame("HDL_32E", *[f[k] for k in FIELDS]); b = R.ref_frame("HDL_32E", *[f[k] for k in FIELDS], double_libm=True)
        assert (a["label"] != b["label"]).any()


def test_boundaries_and_hot_cell(R):
    sp = R.sensor("HDL_32E")
    r, _ = check_frame(R, "HDL_32E", cases.boundary_frame(sp), what="boundaries")
    assert r["multi"].any()
    r, _ = check_frame(R, "HDL_32E", cases.hot_cell_frame(sp), what="hot cell")
    assert (r["single"] > 0).sum() == 1                                    # every point in one cell
    sp = R.sensor("HDL_64E")
    check_frame(R, "HDL_64E", cases.hot_cell_frame(sp), what="hot cell 64")


def test_belonging_grid_and_distance(R):
    rng = np.random.default_rng(0)
    import ctypes as C
    for x, y in np.r_[rng.uniform(-90, 90, (300, 2)), [[-75, -50], [-75.00001, 50], [75, 49.99999], [1e9, -1e9], [-1.0, -1.0], [0.99999994, 1.9999999]]]:
        x = np.float32(x); y = np.float32(y)
        nx = np.float32(np.float64(x) + 75.0); ny = np.float32(np.float64(y) + 50.0)
        want = (int(np.clip(np.floor(np.float64(nx) / 2.0), 0, 74)), int(np.clip(np.floor(np.float64(ny) / 2.0), 0, 49)))
        assert R.ref_belonging_grid(float(x), float(y)) == want
    L = R.ref_bevgen_lib()
    for _ in range(200):
        a = rng.normal(0, 50, 3).astype(np.float32); b = rng.normal(0, 50, 3).astype(np.float32)
        d = (a - b).astype(np.float32)
        want = np.sqrt(np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])))
        got = L.ref_get_distance(a.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.float32(got) == np.float32(want)


@pytest.mark.parametrize("K,seed,step", [(100, 7, 2.0), (400, 21, 2.0), (1500, 8, 2.0), (60, 3, 9.0), (1, 9, 2.0), (2, 10, 2.0), (30, 4, 0.1)])
def test_major_frames_and_labels(R, synth, capfd, K, seed, step):
    """selectMajorFrames / getKeyFrameLabel of the reference's source (real KD-tree, real weights) == oracle, bit for bit;
    K=30 step 0.1 is the M = 1 edge (FLT_MAX second neighbour, w1 written over w0)."""
    xyz = synth.make_poses(K, seed=seed, step=step)
    mi, lab = R.ref_select_and_label(xyz)
    omi, _ = R.select_major(xyz)
    assert np.array_equal(mi, omi)
    olab, _, _ = R.labels(xyz, omi)
    assert lab.shape == olab.shape and np.array_equal(lab.view(np.uint32), olab.view(np.uint32))
    if (K, step) == (30, 0.1):
        assert len(mi) == 1


def test_pose_reader_and_label_writer(R, synth, tmp_path, capfd):
    """readKeyframePose: 16-token rows, `break` at the first row with another token count (:415-419); saveLabels text."""
    xyz = synth.make_poses(12, seed=2, step=9.0)
    lines = synth.pose_csv_lines(xyz)
    p = str(tmp_path / "keyframe_pose.csv")
    open(p, "w").write("\n".join(lines) + "\n")
    got = R.ref_read_poses(p)
    want = np.array([[np.float32(float(v)) for v in l.split(",")[1:4]] for l in lines], np.float32)
    assert np.array_equal(got, want)
    bad = lines[:5] + [",".join(lines[5].split(",")[:15])] + lines[6:]          # row 5 has 15 tokens: reading stops there
    open(p, "w").write("\n".join(bad) + "\n")
    assert np.array_equal(R.ref_read_poses(p), want[:5])
    open(p, "w").write("  ".join(lines) + "\n\n")                               # tokens are whitespace separated, not line separated
    assert np.array_equal(R.ref_read_poses(p), want)
    lab = np.array([[1, 0, 0.25], [1e-39, 0.333333343, 0.6666667], [123456.7, 1e-5, 0.5]], np.float32)
    q = str(tmp_path / "labels.csv")
    R.ref_save_labels(lab, q)
    assert open(q).read() == "".join("".join("%g," % v for v in row) + "\n" for row in lab)


def test_pcd_listing(R, tmp_path):
    d = tmp_path / "kp"; d.mkdir()
    for n in ("000010.pcd", "000002.pcd", "a.b.pcd", "notes.txt", "x.pcd.bak", "pcd", ".pcd"):
        (d / n).write_text("x")
    want = sorted(str(d / n) for n in ("000010.pcd", "000002.pcd", "a.b.pcd", "pcd", ".pcd"))
    assert R.ref_list_pcd(str(d)) == want                                        # "pcd": substr(npos + 1) is the whole name
    assert R.ref_list_pcd(str(d) + "/") == want


def test_reference_main_on_a_folder(R, synth, tmp_path):
    """The reference's whole main() (:664-771) on a keyframe folder vs what the oracle predicts for every output file."""
    import importlib
    cv2 = pytest.importorskip("cv2")
    pcd = importlib.import_module("pcpt_b200.pcd")
    sensor, n = "HDL_32E", 4
    root = str(tmp_path / "kf")
    os.makedirs(os.path.join(root, "keyframe_point_cloud"))
    frames = [synth.make_frame(sensor, 300 + i) for i in range(n)]
    for i, f in enumerate(frames):
        p = os.path.join(root, "keyframe_point_cloud", "%06d.pcd" % i)
        if i == 1:
            pcd.write_ascii(p, f, fields=("label", "x", "y", "z", "col", "row", "intensity", "t"))
        elif i == 2:
            pcd.write_binary_layout(p, f)
        else:
            pcd.write(p, f)
    xyz = synth.make_poses(n, seed=5, step=9.0)
    open(os.path.join(root, "keyframe_pose.csv"), "w").write("\n".join(synth.pose_csv_lines(xyz)) + "\n")
    rc, out, err = R.ref_main(root, sensor)
    assert rc == 0, err[-2000:]
    sp = R.sensor(sensor)
    assert "Using sensor_type %s, with params: N_SCAN: %d, Horizon_SCAN: %d, GROUND_UPPER_SCAN: %d" % (sensor, sp.n_scan, sp.horizon_scan, sp.ground_upper_scan) in out
    assert "[TIME] Average preprocessing and BEV generation: " in out and out.rstrip().endswith("Done.")
    for i, f in enumerate(frames):
        name = "%06d" % i
        assert "Converting file: %s\n" % name in out
        o = R.frame(sp, *[f[k] for k in FIELDS])
        b = np.fromfile(os.path.join(root, "output_multi_bev", "binary", name + ".bin"), np.uint8)
        assert np.array_equal(b.reshape(24, 224, 224), o["multi"]), name
        for l in range(24):
            img = cv2.imread(os.path.join(root, "output_multi_bev", "image", name, "%02d.png" % l), cv2.IMREAD_UNCHANGED)
            assert img is not None and np.array_equal(img, o["multi"][l]), (name, l)
        assert np.array_equal(cv2.imread(os.path.join(root, "output_single_bev", "image", name + ".png"), cv2.IMREAD_UNCHANGED), o["single"])
        assert open(os.path.join(root, "output_single_bev", "csv", name + ".csv")).read() == \
            "\n".join(", ".join("%3d" % v for v in row) for row in o["single"]) + "\n"
        got, hdr = pcd.read(os.path.join(root, "non_ground_point_cloud", name + ".pcd"))
        assert hdr.encode() == pcd.header(sp.S)
        exp = np.zeros(sp.S, pcd.DTYPE)
        sel = o["owner"] > 0; idx = o["owner"][sel].astype(np.int64) - 1
        for k in ("x", "y", "z", "intensity", "row", "col", "t"):
            exp[k][sel] = f[k][idx]
        exp["label"] = o["label"]
        assert pcd.records(got).tobytes() == exp.tobytes(), name
    pxyz = np.array([[np.float32(float(v)) for v in l.split(",")[1:4]] for l in synth.pose_csv_lines(xyz)], np.float32)
    mi, _ = R.select_major(pxyz)
    lab, _, _ = R.labels(pxyz, mi)
    assert open(os.path.join(root, "keyframe_label.csv")).read() == "".join("".join("%g," % v for v in row) + "\n" for row in lab)
    assert "One-hot label has length: %d" % len(mi) in out


def test_cloud_manip_save_as_mat_and_main(R, tmp_path):
    """CloudManip.cpp's own saveAsMat (:79-109) and main() (:111-141): grid, CSV text ("%.4g"), transform."""
    import importlib
    cv2 = pytest.importorskip("cv2")
    pcd = importlib.import_module("pcpt_b200.pcd")
    rng = np.random.default_rng(5)
    n = 200_000
    hot = rng.random(n) < 0.6
    x = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    y = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    z = rng.uniform(-2, 10, n).astype(np.float32)
    z[:50] = np.nan; x[50:100] = np.inf; z[100:150] = -2.0; x[150:160] = -100.5; y[160:170] = 100.49999
    grid = R.ref_save_as_mat(x, y, z, str(tmp_path / "g.csv"))
    want = R.save_as_mat(x, y, z)
    assert np.array_equal(grid.view(np.uint32), want.view(np.uint32)) and (grid > 0).sum() > 10000
    assert open(tmp_path / "g.csv").read() == "\n".join(", ".join("%.4g" % v for v in row) for row in want) + "\n"
    # the tool: Translation * RotZ(theta) then both grids, CSVs, PNGs and PCDs in the working directory
    f = dict(x=x[200:60200], y=y[200:60200], z=z[200:60200], intensity=rng.random(60000).astype(np.float32), row=np.zeros(60000, np.uint16),
             col=np.zeros(60000, np.uint16), t=np.arange(60000, dtype=np.uint32), label=np.full(60000, -2, np.int16))
    pcd.write(str(tmp_path / "c.pcd"), f)
    rc, out, err = R.ref_cloud_manip_main([str(tmp_path / "c.pcd"), "3.5", "-1.25", "0.2", "37"], cwd=str(tmp_path))
    assert rc == 0, err[-1000:]
    rt, (tx, ty, tz) = R.ref_cloud_manip_matrix(3.5, -1.25, 0.2, 37.0, f["x"], f["y"], f["z"])
    th = np.float32(np.float32(37.0) / np.float32(180.0) * np.pi)
    c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    assert np.array_equal(rt, np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, np.float32(np.float32(1) - c) + c, 0.2], np.float32))
    otx, oty, otz = R.transform(rt, f["x"], f["y"], f["z"])
    for a, b in ((tx, otx), (ty, oty), (tz, otz)):
        assert _same_float(a, b)
    got, _ = pcd.read(str(tmp_path / "c.pcd_output.pcd"))
    assert _same_float(got["x"], otx) and _same_float(got["y"], oty) and _same_float(got["z"], otz) and np.array_equal(got["t"], f["t"])
    gi = R.save_as_mat(f["x"], f["y"], f["z"]); go = R.save_as_mat(otx, oty, otz)
    for name, g in (("c.pcd_input.csv", gi), ("c.pcd_output.csv", go)):
        assert open(tmp_path / name).read() == "\n".join(", ".join("%.4g" % v for v in row) for row in g) + "\n"
        png = cv2.imread(str(tmp_path / (name + ".png")), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(png, np.clip(np.rint(g), 0, 255).astype(np.uint8))


def test_batch_cloud_manip_frame(R, synth, tmp_path):
    """BatchCloudManip.cpp's own getOrderedCloud / markGroundPoints / saveAsMat (HDL-64E constants) == oracle order +
    mark_ground + bvm."""
    sp = R.sensor("HDL_64E")
    for idx in (20, 21):
        f = synth.make_frame("HDL_64E", idx)
        lab, m = R.ref_bcm_frame(*[f[k] for k in FIELDS], out_prefix=str(tmp_path / ("b%d" % idx)))
        oc = R.order(sp, *[f[k] for k in FIELDS])
        olab = R.mark_ground(sp, oc)[0]
        assert np.array_equal(lab, olab)
        want = R.bvm(oc, olab)
        assert np.array_equal(m.view(np.uint32), want.view(np.uint32)) and (want > 0).sum() > 200
        assert open(tmp_path / ("b%d.csv" % idx)).read() == "\n".join(", ".join("%.4g" % v for v in row) for row in want) + "\n"


def test_extract_top_and_flatten(R, synth):
    """the oracle's extractTopAndFlatten == TopPartRegistration.cpp:79-141 compiled from the reference's own source
    (oracle/_ref/libtoppart_ref.so, oracle/ref_top_part_shim.cpp).  With equal heights in one cell only the multiset of a cell's
    picks is defined (std::sort), which the last block checks."""
    if R.ref_top_flatten(np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32), np.ones(1, np.int16)) is None:
        pytest.skip("oracle/_ref/libtoppart_ref.so not built")
    sizes = {}
    for name, x, y, z, lab in cases.top_flatten_cases(R, synth):
        rx, ry = R.ref_top_flatten(x, y, z, lab)
        ox, oy, oi = R.top_flatten(x, y, z, lab)
        assert len(rx) == len(ox), (name, len(rx), len(ox))
        assert np.array_equal(rx.view(np.uint32), ox.view(np.uint32)) and np.array_equal(ry.view(np.uint32), oy.view(np.uint32)), name
        assert np.array_equal(x[oi].view(np.uint32), ox.view(np.uint32)), name
        sizes[name] = len(ox)
    assert sizes["keyframe"] > 1000 and sizes["random 300007"] > 10000 and sizes["empty"] == 0
    # cells along x: -95 and -100 share cell 0 (19 + 25 = 44 points -> round(8.8) = 9); 20 -> 4; 22 -> round(4.4) = 4; 23 -> round(4.6) = 5;
    # 9.999 -> cell 5 (30 -> 6); 10.0 (5.5 rounds away from zero) and 10.001 -> cell 6 (60 -> 12); 99.999 and 100 -> cell 10: outside
    assert sizes["thresholds"] == 9 + 4 + 4 + 5 + 6 + 12, sizes
    # ties: quantised heights - the selected heights per cell agree as multisets, and the count is exact
    rng = np.random.default_rng(5)
    n = 50_000
    x = rng.uniform(-100, 100, n).astype(np.float32); y = rng.uniform(-100, 100, n).astype(np.float32)
    z = (rng.integers(-8, 40, n) * 0.25).astype(np.float32); lab = np.ones(n, np.int16)
    rx, ry = R.ref_top_flatten(x, y, z, lab)
    ox, oy, oi = R.top_flatten(x, y, z, lab)
    assert len(rx) == len(ox)
    zmap = {}
    for xi, yi, zi in zip(x.tolist(), y.tolist(), z.tolist()):
        zmap.setdefault((xi, yi), []).append(zi)
    cell = lambda a, b: (int(np.round(np.float32(a + np.float32(100)) / np.float32(20))), int(np.round(np.float32(b + np.float32(100)) / np.float32(20))))
    def per_cell(px, py):
        d = {}
        for a, b in zip(px.tolist(), py.tolist()):
            d.setdefault(cell(a, b), []).append(max(zmap[(a, b)]))
        return {k: sorted(v) for k, v in d.items()}
    assert per_cell(rx, ry) == per_cell(ox, oy)


# ---- the extractors' projection step: {Mulran,Oxford,Kitti}PointCloudSelect.cpp compiled unmodified -------------
def _extractor(R, dataset, root, x, y, z, double_libm):
    inten = (np.arange(len(x)) % 251).astype(np.float32)
    r = R.ref_extract_point_cloud(dataset, str(root), x, y, z, inten, double_libm=double_libm)
    if r is None:
        pytest.skip("oracle/_ref/lib%sselect_ref.so not built" % dataset)
    return r, inten


@pytest.mark.parametrize("double_libm", [False, True])
def test_mulran_and_oxford_projection(R, tmp_path, double_libm):
    """oracle_project_mulran / _oxford == extractPointCloud of the reference's own translation units (MulranPointCloudSelect.cpp:95-130,
    OxfordPointCloudSelect.cpp:146-224), run on scan files in the datasets' layouts: random clouds plus zeros of both signs, axis
    points, huge / tiny / non-finite coordinates; both overload sets of the unqualified atan2 / sqrt / round."""
    x, y, z = cases.projection_cloud()
    # MulRan reads at most 64 * 1024 points (:110); one below that, the read that hits end-of-file appends one more point (:111-127)
    for n in (65_535, 64 * 1024, 70_000, 1, 0):
        r, inten = _extractor(R, "mulran", tmp_path, x[:n], y[:n], z[:n], double_libm)
        m = min(n, 64 * 1024)
        assert len(r["x"]) == min(n + 1, 64 * 1024), (n, len(r["x"]))
        row, col = R.project_mulran(x[:m], y[:m], double_libm=double_libm)
        assert np.array_equal(r["row"][:m], row) and np.array_equal(r["col"][:m], col), n
        assert _same_float(r["x"][:m], x[:m]) and _same_float(r["y"][:m], y[:m]) and _same_float(r["z"][:m], z[:m])
        assert np.array_equal(r["intensity"][:m], inten[:m]) and set(r["label"][:m].tolist()) <= {-2}
        if m:
            assert row.max() == min(m, 64) - 1
        if m > 60_000:
            assert col.max() == 1024                                          # col == Horizon_SCAN is reachable (:125)
    for n in (len(x), 1000, 0):
        r, inten = _extractor(R, "oxford", tmp_path, x[:n], y[:n], z[:n], double_libm)
        nx, nz, row, col = R.project_oxford(x[:n], y[:n], z[:n], double_libm=double_libm)
        assert len(r["x"]) == n
        assert np.array_equal(r["row"], row) and np.array_equal(r["col"], col), (n, int((r["col"] != col).sum()))
        assert _same_float(r["x"], nx) and _same_float(r["z"], nz) and _same_float(r["y"], y[:n])       # upside-down mount: x, z negated
        assert np.array_equal(r["intensity"], inten) and set(r["label"].tolist()) <= {-2}
        if n > 100_000:
            assert col.max() < 1056 and set(np.unique(row)) == set(range(32))


@pytest.mark.parametrize("double_libm", [False, True])
@pytest.mark.parametrize("seed,kw", cases.KITTI_SCANS)
def test_kitti_ring_detection(R, synth, tmp_path, seed, kw, double_libm):
    """oracle_project_kitti == extractPointCloud of KittiPointCloudSelect.cpp (:156-246): the structured 64 x 2083 cloud it returns
    equals the one rebuilt from the oracle's per-point (row, col) - ring detection with short rings, spurious sign flips, a scan that
    starts below 0 degrees, more than 64 rings; the last writer of a slot wins; intensity -1 / label -2 in written slots."""
    x, y, z = synth.make_kitti_scan(seed, **kw)
    r, _ = _extractor(R, "kitti", tmp_path, x, y, z, double_libm)
    row, col = R.project_kitti(x, y, double_libm=double_libm)
    want = cases.kitti_structured(x, y, z, row, col)
    assert len(r["x"]) == 64 * cases.KITTI_H
    for k in want:
        assert _same_float(r[k], want[k]) if want[k].dtype == np.float32 else np.array_equal(r[k], want[k]), (k, seed)
    assert (want["label"] == -2).sum() > 1000


def test_kitti_full_scan_and_collisions(R, tmp_path):
    """A scan of exactly 64 * 2083 points (the extractor's read limit, :172-173: no read past end-of-file) whose rings revisit columns:
    later points replace earlier ones in their slot."""
    n = 64 * cases.KITTI_H
    k = np.arange(n)
    ring = k // 2083
    az = (((k % 2083) * 0.3461) % 360.0)                                      # two passes over each ring's columns: collisions
    az = np.where(az > 180.0, az - 360.0, az) + 1e-3
    rad = 5.0 + (k % 97) * 0.31
    x = (rad * np.cos(np.deg2rad(az))).astype(np.float32); y = (rad * np.sin(np.deg2rad(az))).astype(np.float32)
    z = (ring * 0.05 - 1.7).astype(np.float32)
    r, _ = _extractor(R, "kitti", tmp_path, x, y, z, False)
    row, col = R.project_kitti(x, y)
    want = cases.kitti_structured(x, y, z, row, col)
    for key in want:
        assert _same_float(r[key], want[key]) if want[key].dtype == np.float32 else np.array_equal(r[key], want[key]), key
    placed = row != 0xFFFF
    assert placed.sum() > (want["label"] == -2).sum() > 1000                  # some slots were written more than once
