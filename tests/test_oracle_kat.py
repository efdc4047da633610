"""Hand-derived known-answer tests pinning the CPU oracle: one per arithmetic quirk of SURVEY §8a.1.
Expected values are worked out by hand from the reference source (file:line in each test), not from the oracle."""
import ctypes

import numpy as np
import pytest

from conftest import FIELDS, cat_frames, oracle_batch

f32 = np.float32


def small_sensor(O, N=16, H=24, G=8, hr=0.5):
    s = O.Sensor(); s.n_scan, s.horizon_scan, s.ground_upper_scan, s.height_res = N, H, G, hr
    return s


def frame_from(points):
    """points: list of (row, col, x, y, z, intensity, label)."""
    a = np.array(points, np.float64).reshape(-1, 7)
    return dict(row=a[:, 0].astype(np.uint16), col=a[:, 1].astype(np.uint16), x=a[:, 2].astype(f32), y=a[:, 3].astype(f32),
                z=a[:, 4].astype(f32), intensity=a[:, 5].astype(f32), label=a[:, 6].astype(np.int16))


def run(O, sp, pts):
    f = frame_from(pts)
    return O.frame(sp, *[f[k] for k in FIELDS]), f


def test_sensor_table(O):
    # src/Utility.cpp:92-124 and substring match :72-89
    for name, exp in (("HDL_32E", (32, 1056, 20, 0.5)), ("HDL_64E", (64, 2083, 50, 0.25)), ("OS1_64", (64, 1024, 31, 1.0)),
                      ("xHDL_64Ey", (64, 2083, 50, 0.25))):
        s = O.sensor(name)
        assert (s.n_scan, s.horizon_scan, s.ground_upper_scan, s.height_res) == exp
    with pytest.raises(ValueError):
        O.sensor("VLP16")


def test_order_last_writer_wins_and_bounds(O):
    # BatchMultiBevGen.cpp:102-116: serial loop, later points overwrite; row>=N or col>=H dropped (col == H from MulRan)
    sp = small_sensor(O)
    pts = [(3, 5, 1, 1, 1, 0.5, -2), (3, 5, 2, 2, 2, 0.5, 7), (16, 0, 9, 9, 9, 0.5, 1), (0, 24, 9, 9, 9, 0.5, 1), (3, 5, 3, 3, 3, 0.5, 9)]
    f = frame_from(pts)
    oc = O.order(sp, *[f[k] for k in FIELDS])
    slot = 3 * 24 + 5
    assert oc["owner"][slot] == 5 and oc["x"][slot] == 3 and oc["label"][slot] == 9
    assert (oc["owner"] > 0).sum() == 1
    # unwritten slots are value-initialised: intensity 0 (NOT -1), label 0 (:98)
    assert oc["intensity"][0] == 0 and oc["label"][0] == 0 and oc["x"][0] == 0


def test_cell_index_rounding(O):
    # :279 x = round((px + 112)/1 + 0.5): v in (-1,0) -> 0; v = -1 -> -1 (out); v = 0 -> 1; v = 222.999 -> 223; v = 223 -> 224 (out)
    sp = small_sensor(O)
    vs = [(-113.0, None), (-112.5, 0), (-112.0, 1), (-111.5, 1), (110.9999, 223), (111.0, None), (0.0, 113)]
    for i, (px, cell) in enumerate(vs):
        out, _ = run(O, sp, [(0, i, px, 0.0, 0.0, -1.0, 5)])   # intensity -1 everywhere else irrelevant: row 0 is above the band
        nz = np.argwhere(out["single"] > 0)
        if cell is None:
            assert len(nz) == 0, px
        else:
            assert nz.tolist() == [[cell, 113]], (px, nz)
            assert out["single"][cell, 113] == 8           # int((0+2)*4) = 8 (:345)
            # x is the matrix ROW, y the column (:289): byte offset layer*50176 + x*224 + y; z=0,HR=.5 -> layer round(0/.5+2)=2
            assert out["multi"][2, cell, 113] == 255 and out["multi"].sum() == 255


def test_layer_ties_and_height_truncation(O):
    sp = small_sensor(O, hr=0.5)
    # layer = round(z/0.5 + 2) half away from zero (:281): z=-0.75 -> 0.5 -> 1 ; z=-1.25 -> -0.5 -> -1 (skipped in multi)
    # z=-1.24 -> -0.48 -> 0 ; z=10.75 -> 23.5 -> 24 (skipped); z=10.74 -> 23.48 -> 23
    cases = [(-0.75, 1), (-1.25, None), (-1.24, 0), (10.75, None), (10.74, 23), (0.1, 2), (0.13, 2), (0.125, 2), (0.126, 2), (0.375, 3)]
    for z, layer in cases:
        out, _ = run(O, sp, [(0, 0, 0.0, 0.0, z, 0.5, 5)])
        got = np.argwhere(out["multi"] > 0)
        if layer is None:
            assert len(got) == 0, z
        else:
            assert got.tolist() == [[layer, 113, 113]], (z, got)
        # single: int((z+2)*4) truncation toward zero, clamped (:345-346) — independent of the layer test (:349)
        h = int((float(f32(f32(z) + f32(2.0)))) * 4.0)
        assert out["single"][113, 113] == min(max(h, 0), 255)
    out, _ = run(O, sp, [(0, 0, 0.0, 0.0, 100.0, 0.5, 5)])
    assert out["single"][113, 113] == 255                     # clamp high; layer 202 skipped but single still written
    out, _ = run(O, sp, [(0, 0, 0.0, 0.0, -2.2, 0.5, 5)])
    assert out["single"].sum() == 0                            # (−0.2*4) = −0.8 -> int 0 ; max(0,.)=0
    out, _ = run(O, sp, [(0, 0, 0.0, 0.0, 6e8, 0.5, 5)])
    assert out["single"][113, 113] == 0                        # 2.4e9 overflows int: cvttsd2si -> INT_MIN -> clamp 0


def test_label_zero_skipped_and_max(O):
    sp = small_sensor(O)
    out, _ = run(O, sp, [(0, 0, 0.0, 0.0, 1.0, 0.5, 0), (0, 1, 0.2, 0.2, 3.0, 0.5, 4), (0, 2, 0.4, 0.4, 2.0, 0.5, 4)])
    assert out["single"][113, 113] == 20 and (out["single"] > 0).sum() == 1     # max of 20 and 16; label 0 point ignored
    assert sorted(np.argwhere(out["multi"] > 0)[:, 0].tolist()) == [6, 8]       # layers round(2/.5+2)=6, round(3/.5+2)=8


def test_sector_index_and_clamp(O):
    # BatchMultiBevGen.h:78-96: floor((x+75)/2) clamped to [0,74], floor((y+50)/2) to [0,49]
    import bevgen_oracle_py as P
    assert P.belonging_grid(f32(0), f32(0)) == (37, 25)
    assert P.belonging_grid(f32(-75.1), f32(-50.1)) == (0, 0)
    assert P.belonging_grid(f32(-0.999), f32(1.999)) == (37, 25)
    assert P.belonging_grid(f32(1.0), f32(2.0)) == (38, 26)
    assert P.belonging_grid(f32(500), f32(500)) == (74, 49)
    assert P.belonging_grid(f32(5e9), f32(np.nan)) == (0, 0)    # floor(x/2) >= 2^31 and NaN: cvttsd2si -> INT_MIN -> clamp 0


def flat_ground_frame(N, H, z=-1.7, r0=4.0, dr=1.5):
    pts = []
    for r in range(N):
        for c in range(H):
            rad = r0 + (N - 1 - r) * dr
            a = 2 * np.pi * c / H
            pts.append((r, c, rad * np.cos(a), rad * np.sin(a), z, 0.5, -2))
    return pts


def test_ground_band_and_row_above(O):
    # flat plane: every pair in the band is ground (angle 0). Band rows N-G..N-1 plus row N-G-1 get gm=1 (:179-182)
    sp = small_sensor(O)                      # N=16, G=8: band rows 8..15, row 7 marked via row-1
    f = frame_from(flat_ground_frame(16, 24))
    oc = O.order(sp, *[f[k] for k in FIELDS])
    lab, gm1, gmf, avg = O.mark_ground(sp, oc)
    assert (gm1[7:] == 1).all() and (gm1[:7] == 0).all()
    assert (lab.reshape(16, 24)[7:] == 0).all() and (lab.reshape(16, 24)[:7] == -2).all()


def test_count_sequence_and_mean(O):
    # :135 count starts at float(0.01); :205 +1 per ground point; :210 mean = sum/count (IEEE float divide)
    sp = small_sensor(O, N=16, H=24, G=8)
    # all 9 rows x 24 cols of ground land in ONE sector (37,25): x,y in [0,1)
    pts = []
    for r in range(16):
        for c in range(24):
            pts.append((r, c, 0.01 * c, 0.5 + 0.001 * r, -1.5 if r >= 7 else 5.0, 0.5, -2))
    f = frame_from(pts)
    oc = O.order(sp, *[f[k] for k in FIELDS])
    lab, gm1, gmf, avg = O.mark_ground(sp, oc)
    n = int((gm1 == 1).sum())
    assert n == 9 * 24
    cnt = f32(0.01)
    s = f32(0)
    for _ in range(n):
        cnt = f32(cnt + f32(1)); s = f32(s + f32(-1.5))
    assert avg[37, 25] == f32(s / cnt)
    assert (np.delete(avg.ravel(), 37 * 50 + 25) == 0).all()     # untouched sectors: 0 / 0.01 = 0
    # sequence check of the count itself: 1.00999999, 2.00999999, 3.00999999, 4.01000023
    c = f32(0.01); seq = []
    for _ in range(4):
        c = f32(c + f32(1)); seq.append(float(c))
    assert seq == [float(f32(1.01)), float(f32(2.01)), float(f32(3.01)), 4.010000228881836]


def test_own_sector_not_tested_and_neighbour_order(O):
    # loop 3 (:216-250) tests the FOUR neighbours only — a tall "ground" point is kept ground if only its own sector is low
    sp = small_sensor(O, N=16, H=24, G=8)
    pts = flat_ground_frame(16, 24, z=-1.7, r0=30.0, dr=0.01)   # far ring: all points of a column fall into few sectors
    f = frame_from(pts)
    oc = O.order(sp, *[f[k] for k in FIELDS])
    lab, gm1, gmf, avg = O.mark_ground(sp, oc)
    # neighbours of occupied sectors are mostly empty (avg 0): z - 0 = -1.7 > 0.30 is false -> stays ground
    assert (gmf[7:] == 1).all()
    # raise every point to z = +0.31: now z - avg(empty neighbour = 0) = 0.31 > 0.30 -> cleared unless all 4 neighbours hold ground
    pts2 = [(r, c, x, y, 0.31, i, l) for (r, c, x, y, z, i, l) in pts]
    f2 = frame_from(pts2)
    oc2 = O.order(sp, *[f2[k] for k in FIELDS])
    lab2, gm12, gmf2, avg2 = O.mark_ground(sp, oc2)
    assert (gm12[7:] == 1).all() and (gmf2 == 0).sum() > 0
    assert (lab2.reshape(16, 24)[gmf2 == 0] == -2).all()          # label untouched when not ground (:246-248)


def test_threshold_030_is_a_double_compare(O):
    # (double)(float z - float avg) > 0.30: 0.3f (=0.300000012) passes, the float below (0.29999998) does not
    sp = small_sensor(O, N=16, H=24, G=8)
    for dz, cleared in ((f32(0.3), True), (np.nextafter(f32(0.3), f32(0)), False)):
        pts = [(r, c, 0.01 * c, 0.5, float(dz), 0.5, -2) for r in range(16) for c in range(24)]
        f = frame_from(pts)
        oc = O.order(sp, *[f[k] for k in FIELDS])
        lab, gm1, gmf, avg = O.mark_ground(sp, oc)
        assert avg[37, 25] != 0 and avg[36, 25] == 0               # neighbour (-1,0) is empty: avg 0
        assert ((gmf[7:] == 0).all()) == cleared


def test_intensity_substitution_chain(O):
    # :146-160: upper = (r-1,c); if I==-1 -> (r-1,(c+2)%H); if still -1 -> (r-1,(c-2)%H) [negative for c<2: tail of row r-2];
    # if still -1 and r>=2 -> (r-2,c).  :162 invalid if lower or the final upper is -1.
    sp = small_sensor(O, N=16, H=24, G=8)
    base = flat_ground_frame(16, 24)
    def variant(mods):
        pts = [list(p) for p in base]
        for (r, c), kv in mods.items():
            for k, v in kv.items():
                pts[r * 24 + c][{"x": 2, "y": 3, "z": 4, "i": 5}[k]] = v
        f = frame_from([tuple(p) for p in pts])
        oc = O.order(sp, *[f[k] for k in FIELDS])
        return O.mark_ground(sp, oc)
    # (a) direct upper (14,5) flagged; (14,7) is a wall point far above -> pair (15,5) uses it -> NOT ground
    lab, gm1, _, _ = variant({(14, 5): {"i": -1}, (14, 7): {"z": 30.0}})
    assert gm1[15, 5] == 0
    # (b) both (14,5),(14,7) flagged -> falls to (14,3), a wall -> not ground; with (14,3) flat -> ground
    lab, gm1, _, _ = variant({(14, 5): {"i": -1}, (14, 7): {"i": -1}, (14, 3): {"z": 30.0}})
    assert gm1[15, 5] == 0
    # (c) all three flagged -> (13,5) wall -> not ground
    lab, gm1, _, _ = variant({(14, 5): {"i": -1}, (14, 7): {"i": -1}, (14, 3): {"i": -1}, (13, 5): {"z": 30.0}})
    assert gm1[15, 5] == 0
    # (d) everything flagged incl (13,5) -> invalid: gm = -1 (:165)
    lab, gm1, _, _ = variant({(14, 5): {"i": -1}, (14, 7): {"i": -1}, (14, 3): {"i": -1}, (13, 5): {"i": -1}})
    assert gm1[15, 5] == -1
    # (e) col 1: (c-2)%H = -1 -> index (r-1)*H - 1 = slot (13, 23): make THAT a wall
    lab, gm1, _, _ = variant({(14, 1): {"i": -1}, (14, 3): {"i": -1}, (13, 23): {"z": 30.0}})
    assert gm1[15, 1] == 0
    lab, gm1, _, _ = variant({(14, 1): {"i": -1}, (14, 3): {"i": -1}, (14, 23): {"z": 30.0}})
    assert gm1[15, 1] == 1                                       # (14,23) is NOT what C++ % selects
    # (f) invalid overwrites the 1 written by the row below (rows descend): lower (14,5) flagged
    lab, gm1, _, _ = variant({(14, 5): {"i": -1}})
    assert gm1[14, 5] == -1 and gm1[13, 5] == 1


def test_angle_overload_and_threshold(O):
    # :173,:179: float atan2f/sqrtf; ground iff fabsf(angle) <= 10.0f with angle = float(double(atan2f)*180/M_PI)
    t = np.array([0x3e32b8c2], np.uint32).view(f32)[0]            # a_max
    assert f32(np.float64(t) * 180.0 / np.pi) <= f32(10.0) < f32(np.float64(np.nextafter(t, f32(1))) * 180.0 / np.pi)
    lib = O.lib()
    assert lib.oracle_angle_deg(0.0, 0.0, 0.0) == 0.0              # atan2(0,0)=0: two empty slots are "ground"
    assert abs(lib.oracle_angle_deg(1.0, 1.0, 0.0) - 45.0) < 1e-5
    assert lib.oracle_angle_deg(-1.0, 0.0, 0.0) == -90.0


def test_labels_weights_by_hand(O):
    # :623-630 with d0^2 = 1, d1^2 = 4: w0 = 1/(1+1e-5), w1 = 1/(4+1e-5) then normalised in float
    xyz = np.array([[0, 0, 0], [25, 0, 0], [1, 0, 0]], f32)
    mi, ov = O.select_major(xyz)
    assert mi.tolist() == [0, 1] and ov.tolist() == [-1, -1, 0]   # 24 m from the last major (25,0,0): step 2 finds major 0 at 1 m
    lab, nn, w = O.labels(xyz, mi)
    assert lab[0].tolist() == [1.0, 0.0] and lab[1].tolist() == [0.0, 1.0]
    w0 = f32(1.0 / (1.0 + 1e-5)); w1 = f32(1.0 / (576.0 + 1e-5)); s = f32(w0 + w1)
    assert lab[2].tolist() == [float(f32(w0 / s)), float(f32(w1 / s))]
    # M == 1 edge (:629-630): d1 = FLT_MAX, cand1 = 0 -> w1 overwrites w0 at index 0
    xyz = np.array([[0, 0, 0], [1, 0, 0]], f32)
    mi, _ = O.select_major(xyz)
    lab, nn, w = O.labels(xyz, mi)
    assert mi.tolist() == [0] and lab[0, 0] == 1.0
    fmax = float(np.finfo(f32).max)
    e0 = f32(1.0 / (1.0 + 1e-5)); e1 = f32(1.0 / (fmax + 1e-5))
    assert lab[1, 0] == f32(e1 / f32(e0 + e1)) and 0 < lab[1, 0] < 1e-38


def test_major_selection_rule(O):
    # :527-558: frame is major iff dist to last major >= 20 AND squared dist to nearest earlier major >= 400
    xyz = np.array([[0, 0, 0], [19.99, 0, 0], [20, 0, 0], [40, 0, 0], [20, 15, 0], [0.5, 0, 0]], f32)
    mi, ov = O.select_major(xyz)
    # idx4: 25 m from last major (40,0) but 15 m from major (20,0) -> overlap with major #1; idx5: 39.5 from last, 0.5 from major #0
    assert mi.tolist() == [0, 2, 3] and ov.tolist() == [-1, -2, -1, -1, 1, 0]


def test_cv2_divide_and_u8_semantics():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    a = rng.normal(0, 100, (75, 50)).astype(f32); b = (rng.random((75, 50)) * 50 + 0.01).astype(f32)
    assert np.array_equal(cv2.divide(a, b), (a / b).astype(f32))           # :210 MatExpr '/' == IEEE single divide
    # PNG round trip is lossless for 8-bit gray (parity on decoded pixels)
    img = rng.integers(0, 256, (224, 224), dtype=np.uint8)
    ok, buf = cv2.imencode(".png", img)
    assert ok and np.array_equal(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED), img)


def test_transform_and_save_as_mat(O):
    # CloudManip.cpp:128 PCL SSE order m0*x + (m1*y + (m2*z + t)); :79-95 201x201 max of z+2, strict '>' vs init 0
    rt = np.array([0.5, 0.25, 0.125, 1.0, 0, 1, 0, 0, 0, 0, 1, 0], f32)
    x = np.array([1e8], f32); y = np.array([3.0], f32); z = np.array([7.0], f32)
    tx, ty, tz = O.transform(rt, x, y, z)
    want = f32(f32(x[0] * f32(0.5)) + f32(f32(y[0] * f32(0.25)) + f32(f32(z[0] * f32(0.125)) + f32(1.0))))
    assert tx[0] == want and ty[0] == 3.0 and tz[0] == 7.0
    m = O.save_as_mat(np.array([0, 0, -100.5, 99.99, 100.0], f32), np.array([0, 0, 0, 0, 0], f32), np.array([-2, 1, 5, 5, 5], f32))
    assert m[101, 101] == 3.0 and m[0, 101] == 7.0 and m[200, 101] == 7.0 and (m > 0).sum() == 3


def test_bvm_label_filter_and_strict_max(O):
    """BatchCloudManip.cpp:201-226: label == 0 skipped (:214), cell starts at 0 and takes z + 2 only if strictly greater,
    x = round(px + 100 + 0.5) in [0, 201)."""
    x = np.array([0.2, 0.2, 0.2, -100.9, 99.99, 100.0, 5.0], np.float32)
    y = np.array([0.7, 0.7, 0.7, -100.9, 99.99, 0.0, 5.0], np.float32)
    z = np.array([1.0, 3.0, 9.0, 0.5, 0.25, 1.0, -2.5], np.float32)
    lab = np.array([-2, 7, 0, 1, 1, 1, 1], np.int16)       # the 9.0 point is ground (label 0) and must not count
    m = O.bvm(dict(x=x, y=y, z=z), lab)
    assert m.shape == (201, 201)
    assert m[101, 101] == np.float32(5.0)                   # max(1+2, 3+2); x = floor(100.2)+1 = 101
    assert m[0, 0] == np.float32(2.5)                       # v = -0.9 -> round(-0.4) = -0 -> cell 0
    assert m[200, 200] == np.float32(2.25)                  # v = 199.99 -> 200
    assert m[105 + 1 - 0, 105 + 1 - 0] == 0                 # z + 2 = -0.5 is not > 0: never stored
    assert (m > 0).sum() == 3                               # px = 100.0 -> x = 201: out of range


def test_projection_known_answers(O):
    """Projection step of the extractors: MulranPointCloudSelect.cpp:112-126, OxfordPointCloudSelect.cpp:201-219."""
    x = np.array([1, 0, -1, 0, 1, 1], np.float32); y = np.array([0, 1, 0, -1, -1e-7, np.nan], np.float32)
    row, col = O.project_mulran(x, y)
    assert row.tolist() == [0, 1, 2, 3, 4, 5]                          # k % 64
    # 0, 90, 180 deg; -90 -> 270; a tiny negative angle wraps to 360.0f -> col == 1024 (== Horizon_SCAN, dropped later); NaN -> 0
    assert col.tolist() == [0, 256, 512, 768, 1024, 0]
    r2, _ = O.project_mulran(np.ones(130, np.float32), np.zeros(130, np.float32))
    assert r2[63] == 63 and r2[64] == 0 and r2[129] == 1
    # Oxford: x, z negated; elevation 0 -> round(10.67 / 1.3335) = 8; +10.67 deg -> row 0; below -30.67 deg clamps to 31
    xo = np.array([-1, -1, -1, 0], np.float32); yo = np.array([0, 0, 0, 1], np.float32)
    zo = np.array([0, -np.tan(np.deg2rad(10.67)), 5, 0], np.float32)
    nx, nz, row, col = O.project_oxford(xo, yo, zo)
    assert nx.tolist() == [1, 1, 1, 0] and nz[2] == -5                 # (the 4th x is -0.0 == 0)
    assert row.tolist() == [8, 0, 31, 8]
    assert col.tolist() == [0, 0, 0, 264]                              # 90 deg of 1056 columns
    # col wrap: an azimuth that rounds to 1056 comes back as 0 (:217)
    _, _, _, c = O.project_oxford(np.array([-1], np.float32), np.array([-1e-7], np.float32), np.array([0], np.float32))
    assert c.tolist() == [0]


def test_kitti_ring_detection_rule(O):
    """KittiPointCloudSelect.cpp:199-221: a crossing (az[i-1] <= 0 < az[i]) opens a new ring only when the current ring
    holds more than 2083 * 0.60f = 1249.8 points; point 0 is never placed; a scan that starts at az <= 0 has no ring
    until its first crossing."""
    def scan(n_first, n_rest, first_az=10.0):
        az = np.concatenate([np.full(n_first, first_az), [-10.0], np.full(n_rest, 10.0)])
        return np.cos(np.deg2rad(az)).astype(np.float32), np.sin(np.deg2rad(az)).astype(np.float32)
    x, y = scan(1300, 50)                     # crossing at i = 1301 with num = 1300 > 1249.8 -> ring 1
    row, col = O.project_kitti(x, y)
    assert row[0] == 0xFFFF and col[0] == 0xFFFF
    assert set(row[1:1301]) == {0} and set(row[1301:]) == {1}
    assert col[1] == round(10.0 / (360.0 / 2083)) and col[1300] == round(350.0 / (360.0 / 2083))
    x, y = scan(1250, 50)                     # num = 1250 at the crossing: 1250 > 1249.8 -> accepted
    assert set(O.project_kitti(x, y)[0][1251:]) == {1}
    x, y = scan(1249, 50)                     # num = 1249: ignored, the ring index stays 0
    assert set(O.project_kitti(x, y)[0][1:]) == {0}
    x, y = scan(5, 50, first_az=-10.0)        # starts below 0: nothing is placed before the first crossing, which is always taken
    row, _ = O.project_kitti(x, y)
    assert set(row[:6]) == {0xFFFF} and set(row[6:]) == {0}


def test_top_flatten_quota_and_order(O):
    """extractTopAndFlatten (TopPartRegistration.cpp:79-141): cell = round((p + 100) / 20), cells with < 20 points give
    nothing, others their round(0.2f * count) highest points, highest first; label 0 skipped; cells in x-major order."""
    # cell (5, 5): 30 points around the origin with heights 0..29 -> 6 highest = 29..24; five of them are label 0 -> 25 points -> 5
    n = 30
    x = np.full(n, 1.0, np.float32); y = np.full(n, -2.0, np.float32); z = np.arange(n, dtype=np.float32)
    lab = np.full(n, -2, np.int16)
    ox, oy, oi = O.top_flatten(x, y, z, lab)
    assert oi.tolist() == [29, 28, 27, 26, 25, 24] and set(ox) == {1.0} and set(oy) == {-2.0}
    lab[[29, 27, 0, 1, 2]] = 0
    assert O.top_flatten(x, y, z, lab)[2].tolist() == [28, 26, 25, 24, 23]
    # 19 points: below MIN_GRID_POINTS_SIZE
    assert len(O.top_flatten(x[:19], y[:19], z[:19], np.full(19, 1, np.int16))[0]) == 0
    # rounding of the cell index: x = -90.1 -> round(0.495) = 0, x = -89.9 -> round(0.505) = 1; x = 90 -> round(9.5) = 10: outside
    xs = np.concatenate([np.full(20, -90.1), np.full(20, -89.9), np.full(20, 90.0)]).astype(np.float32)
    ys = np.zeros(60, np.float32); zs = np.tile(np.arange(20, dtype=np.float32), 3); ls = np.ones(60, np.int16)
    ox, oy, oi = O.top_flatten(xs, ys, zs, ls)
    assert oi.tolist() == [19, 18, 17, 16, 39, 38, 37, 36]          # cell (0,5) then (1,5); the x = 90 points fall outside
    # equal heights keep their input order (the stable choice among what std::sort may return)
    zt = np.zeros(40, np.float32); zt[[7, 30]] = 5.0
    assert O.top_flatten(np.zeros(40, np.float32), np.zeros(40, np.float32), zt, np.ones(40, np.int16))[2].tolist() == [7, 30, 0, 1, 2, 3, 4, 5]
