"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for owner / label / single / multi (integer & byte work); labels within 1e-5
relative (north_star).  Run with `pytest -m gpu` on the B200 box."""
import numpy as np
import pytest

from conftest import FIELDS, assert_same, cat_frames, oracle_batch
import cases
from cases import rand_frame as _rand_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gens(pkg):
    cache = {}

    def get(sensor, **kw):
        key = (sensor, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = pkg.BevGen(sensor, device=0, **kw)
        return cache[key]
    yield get
    for g in cache.values():
        g.close()


@pytest.mark.parametrize("sensor,nf", [("HDL_64E", 6), ("OS1_64", 6), ("HDL_32E", 8)])
def test_synthetic_frames_bit_exact(gens, synth, O, sensor, nf):
    batch = synth.make_batch(sensor, nf)
    g = gens(sensor, max_frames_per_batch=4)          # nf > 4: at least two waves, both lanes of the host pipeline
    out = g.process_host(batch)
    assert_same(out, oracle_batch(O, sensor, batch), sensor)
    assert g.kernel_launches() > 0


def test_sector_mean_sweep_fallback(pkg, synth, O, monkeypatch):
    """Frames with more segments than the segment-form sector-mean kernel holds are swept by k_sector_mean; force that
    with a tiny capacity and compare both routes with the oracle."""
    batch = synth.make_batch("HDL_32E", 3, first=40)
    ref = oracle_batch(O, "HDL_32E", batch)
    for cap in ("0", "700", "4096"):          # all frames swept / some swept (~700 segments per frame) / none swept
        monkeypatch.setenv("BEVGEN_SEG_CAP", cap)
        g = pkg.BevGen("HDL_32E", device=0, max_frames_per_batch=4)
        try:
            assert_same(g.process_host(batch), ref, "seg_cap=" + cap)
        finally:
            g.close()


def test_frames_above_the_first_segment_capacity(pkg, synth, O):
    """Synthetic HDL_64E keyframe 2030 has 4195 single-sector segments, more than the two-CTAs-per-SM segment build holds (4096):
    it must go through the second, larger build (not the one-warp sweep) and still be bit-exact; mixed with ordinary frames."""
    frames = [synth.make_frame("HDL_64E", i) for i in (2029, 2030, 2031, 2030)]
    offs = np.zeros(len(frames) + 1, np.int64); offs[1:] = np.cumsum([len(f["x"]) for f in frames])
    batch = {k: np.concatenate([f[k] for f in frames]) for k in FIELDS}; batch["offsets"] = offs
    ref = oracle_batch(O, "HDL_64E", batch)
    g = pkg.BevGen("HDL_64E", device=0, max_frames_per_batch=4)
    try:
        assert_same(g.process_host(batch), ref, "frame 2030 (4195 segments)")
    finally:
        g.close()


def test_large_range_image_global_claim_path(pkg, O):
    """A range image too large for the shared-memory ordering kernel (S = 128 x 8192 > 640 k slots) takes the
    global-memory claim path (k_order_claim / k_order_fill / k_winner_bits); same results expected."""
    p = pkg.sensor_params("HDL_64E")
    p.n_scan, p.horizon_scan, p.ground_upper_scan, p.height_res = 128, 8192, 100, 0.25
    sp = O.Sensor(); sp.n_scan, sp.horizon_scan, sp.ground_upper_scan, sp.height_res = 128, 8192, 100, 0.25
    rng = np.random.default_rng(99)
    frames = [_rand_frame(rng, 128, 8192, n, spread=40.0) for n in (300_000, 70_001, 0)]
    # a structured patch so that ground marking has something to do: a flat disc of points over many rows / columns
    n = 200_000
    rows = rng.integers(60, 128, n); cols = rng.integers(0, 8192, n)
    ang = cols / 8192.0 * 2 * np.pi; rad = 3.0 + (127 - rows) * 0.3
    frames.append(dict(x=(rad * np.cos(ang)).astype(np.float32), y=(rad * np.sin(ang)).astype(np.float32),
                       z=(-1.7 + rng.normal(0, 0.02, n)).astype(np.float32), intensity=rng.random(n).astype(np.float32),
                       row=rows.astype(np.uint16), col=cols.astype(np.uint16), label=np.full(n, -2, np.int16)))
    batch = cat_frames(frames)
    g = pkg.BevGen(p, device=0, max_frames_per_batch=2, max_points_per_frame=300_000)
    try:
        out = g.process_host(batch)
    finally:
        g.close()
    ref = oracle_batch(O, sp, batch)
    assert (ref["label"][3][ref["owner"][3] > 0] == 0).mean() > 0.3      # the disc really is ground
    assert_same(out, ref, "large S")


def test_kitti_quirk_all_intensity_minus_one(gens, synth, O):
    batch = synth.make_batch("HDL_64E", 2, first=100, kitti_quirk=True)
    out = gens("HDL_64E", max_frames_per_batch=4).process_host(batch)
    ref = oracle_batch(O, "HDL_64E", batch)
    assert_same(out, ref, "kitti")
    # nearly every pair is invalid (valid points all carry intensity -1) => ground removal is almost a no-op
    assert (ref["label"][ref["owner"] > 0] == -2).mean() > 0.95


@pytest.mark.parametrize("sensor", ["HDL_32E", "OS1_64", "HDL_64E"])
def test_random_unstructured_frames(gens, O, sensor):
    """Uniform random points: heavy slot collisions (last writer wins), out-of-range row/col, mixed labels incl. 0,
    many -1 intensities (substitution chain incl. the negative (col-2)%H wrap), points outside the BEV range."""
    sp = O.sensor(sensor)
    frames = cases.random_unstructured_frames(sp)
    batch = cat_frames(frames)
    g = gens(sensor, max_frames_per_batch=4, max_points_per_frame=sp.S * 2)
    assert_same(g.process_host(batch), oracle_batch(O, sensor, batch), sensor)


def test_ground_plane_borderline_angles(gens, O):
    """Organised frame whose vertical neighbours sit within a few ulps of the 10-degree threshold, plus special
    values (zero-length pairs, huge, inf, nan).  The decision must match glibc's float atan2f exactly."""
    sensor = "HDL_32E"
    sp = O.sensor(sensor)
    f, f2 = cases.borderline_frames(sp)
    batch = cat_frames([f, f2])
    g = gens(sensor, max_frames_per_batch=4, max_points_per_frame=sp.S * 2)
    ref = oracle_batch(O, sensor, batch)
    assert_same(g.process_host(batch), ref, "borderline")
    # the construction really is borderline: float- and double-libm oracles disagree somewhere
    refd = oracle_batch(O, sensor, batch, double_libm=True)
    assert (refd["label"] != ref["label"]).any()


def test_atan2f_bit_exact(gens, O):
    """Device port of glibc's atan2f vs the host libm, bit for bit (NaN payloads aside)."""
    g = gens("HDL_32E", max_frames_per_batch=4, max_points_per_frame=33792 * 2)
    rng = np.random.default_rng(3)
    n = 4_000_000
    bits = rng.integers(0, 2 ** 32, 2 * n, dtype=np.uint64).astype(np.uint32)
    y = bits[:n].view(np.float32).copy(); x = bits[n:].view(np.float32).copy()
    y[: n // 2] = rng.normal(0, 2, n // 2).astype(np.float32)
    x[: n // 2] = np.abs(rng.normal(0, 10, n // 2)).astype(np.float32)
    t = np.float32(np.tan(0.17453292)); q = slice(n // 4, n // 2)
    y[q] = (x[q].astype(np.float64) * t * (1 + rng.integers(-50, 51, n // 4) * 6e-8)).astype(np.float32)
    sp = [0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, 2.0 ** 25, 2.0 ** 26, 0.4375, 0.6875, 1.1875, 2.4375]
    sy, sx = np.meshgrid(np.array(sp, np.float32), np.array(sp, np.float32))
    y[: sy.size] = sy.ravel(); x[: sx.size] = sx.ravel()
    got = g.debug_atan2f(y, x)
    import ctypes
    lm = ctypes.CDLL("libm.so.6"); lm.atan2f.restype = ctypes.c_float; lm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    # vectorised host reference through the oracle's libm binding on a subsample + numpy-free full check via ctypes loop on 200k
    sel = np.concatenate([np.arange(sy.size), rng.choice(n, 200_000, replace=False)])
    want = np.array([lm.atan2f(float(a), float(b)) for a, b in zip(y[sel], x[sel])], np.float32)
    gg = got[sel]
    same = (gg.view(np.uint32) == want.view(np.uint32)) | (np.isnan(gg) & np.isnan(want))
    assert same.all(), (y[sel][~same][:5], x[sel][~same][:5], gg[~same][:5], want[~same][:5])


def test_cell_index_boundaries(gens, O):
    """Cell / layer / height rounding at the boundaries listed in SURVEY §8a.1-B."""
    sensor = "HDL_32E"
    sp = O.sensor(sensor)
    f = cases.boundary_frame(sp)
    batch = cat_frames([f])
    g = gens(sensor, max_frames_per_batch=4, max_points_per_frame=sp.S * 2)
    ref = oracle_batch(O, sensor, batch)
    assert ref["multi"].any() and ref["single"].any()
    assert_same(g.process_host(batch), ref, "boundaries")


def test_submit_collect_and_device_path(gens, synth, O, pkg):
    import torch
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 5, first=50)
    ref = oracle_batch(O, sensor, batch)
    g = gens(sensor, max_frames_per_batch=4, max_points_per_frame=33792 * 2)
    offs = batch["offsets"]
    fr = lambda f: {k: batch[k][offs[f]:offs[f + 1]] for k in FIELDS}
    chk = lambda f, o: assert_same({k: v[None] for k, v in o.items()}, {k: v[f:f + 1] for k, v in ref.items()}, "submit/collect %d" % f)
    for f in range(4):                               # ring holds max_frames_per_batch (4) frames
        g.submit(100 + f, fr(f))
    with pytest.raises(pkg.BevgenError, match="ring full"):
        g.submit(104, fr(4))
    chk(2, g.collect(102, fr(2)))                    # out-of-order collect frees a slot
    g.submit(104, fr(4))
    for f in (4, 0, 1, 3):
        chk(f, g.collect(100 + f, fr(f)))
    with pytest.raises(pkg.BevgenError):
        g.collect(999)
    # device-resident path: torch only owns the device memory
    dev = torch.device("cuda:0")
    din = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).to(dev) for k in FIELDS}
    S = g.S
    dout = dict(label=torch.empty((5, S), dtype=torch.int16, device=dev),
                winner=torch.zeros(pkg.winner_words(int(offs[-1]), 5), dtype=torch.int32, device=dev),
                single=torch.empty((5, 224 * 224), dtype=torch.uint8, device=dev),
                multi=torch.empty((5, 24 * 224 * 224), dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()
    g.process_device(5, offs, {k: v.data_ptr() for k, v in din.items()}, {k: v.data_ptr() for k, v in dout.items()})
    g.sync()
    own = pkg.owner_from_winner(dout["winner"].cpu().numpy().view(np.uint32), offs, np.ascontiguousarray(batch["row"], np.uint16),
                                np.ascontiguousarray(batch["col"], np.uint16), g.params.horizon_scan, S)
    got = dict(label=dout["label"].cpu().numpy(), owner=own,
               single=dout["single"].cpu().numpy().reshape(5, 224, 224), multi=dout["multi"].cpu().numpy().reshape(5, 24, 224, 224))
    assert_same(got, ref, "device path")


def test_rigid_transform_fused(pkg, synth, O):
    """has_transform: every point goes through R|t (pcl::transformPointCloud op order) before ordering."""
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 2, first=70)
    th = np.float32(np.float32(37.0) / np.float32(180.0) * np.pi)
    c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
    g = pkg.BevGen(sensor, rt=rt, max_frames_per_batch=4)
    out = g.process_host(batch)
    tx, ty, tz = O.transform(rt, batch["x"], batch["y"], batch["z"])
    b2 = dict(batch); b2["x"], b2["y"], b2["z"] = tx, ty, tz
    assert_same(out, oracle_batch(O, sensor, b2), "transform")
    g.close()


def test_labels_and_major_frames(gens, synth, O):
    g = gens("HDL_32E", max_frames_per_batch=4, max_points_per_frame=33792 * 2)
    for K, seed in ((100, 7), (1500, 8), (1, 9), (2, 10)):
        xyz = synth.make_poses(K, seed=seed)
        mi, ov = g.select_major(xyz)
        omi, oov = O.select_major(xyz)
        assert np.array_equal(mi, omi) and np.array_equal(ov, oov), K
        lab, nn, w = g.labels(xyz, mi)
        olab, onn, ow = O.labels(xyz, omi)
        assert np.array_equal(nn, onn), K
        np.testing.assert_allclose(lab, olab, rtol=1e-5, atol=0)      # tolerance stated by north_star
        np.testing.assert_array_equal(lab, olab)                      # and in fact bit-exact
        # row split as the multi-GPU label stage does it
        h = K // 2
        a, _, _ = g.labels(xyz, mi, 0, h); b, _, _ = g.labels(xyz, mi, h, K)
        np.testing.assert_array_equal(np.concatenate([a, b]), olab)


def test_cloud_manip_contention(gens, O):
    """BASELINE config #5 shape (scaled): 60 % of the points in a 3 m blob around the origin (hot cells)."""
    g = gens("HDL_32E", max_frames_per_batch=4, max_points_per_frame=33792 * 2)
    rng = np.random.default_rng(5)
    n = 300_000
    hot = rng.random(n) < 0.6
    x = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    y = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    z = rng.uniform(-2, 10, n).astype(np.float32)
    z[:50] = np.nan; x[50:100] = np.inf; z[100:150] = -2.0
    th = np.float32(np.float32(37.0) / np.float32(180.0) * np.pi)
    c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
    (tx, ty, tz), bi, bo = g.cloud_manip(rt, x, y, z)
    otx, oty, otz = O.transform(rt, x, y, z)
    for a, b in ((tx, otx), (ty, oty), (tz, otz)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    np.testing.assert_array_equal(bi, O.save_as_mat(x, y, z))
    np.testing.assert_array_equal(bo, O.save_as_mat(otx, oty, otz))


def test_full_size_properties(gens, synth):
    """BASELINE-size properties that need no oracle: idempotence (same input twice => identical bytes), owner is a
    valid last-writer map, multi occupancy implies single coverage, labels only change to 0."""
    sensor = "HDL_64E"
    batch = synth.make_batch(sensor, 3, first=200)
    g = gens(sensor, max_frames_per_batch=4)
    a = g.process_host(batch); b = g.process_host(batch)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    offs = batch["offsets"]
    H = g.params.horizon_scan
    for f in range(3):
        own = a["owner"][f]
        idx = own[own > 0].astype(np.int64) - 1
        sl = np.nonzero(own > 0)[0]
        r = batch["row"][offs[f]:offs[f + 1]].astype(np.int64); c = batch["col"][offs[f]:offs[f + 1]].astype(np.int64)
        assert np.array_equal(r[idx] * H + c[idx], sl)
        last = np.full(g.S, -1, np.int64); np.maximum.at(last, r * H + c, np.arange(len(r)))
        assert np.array_equal(last[sl], idx)
        lab = a["label"][f]
        assert set(np.unique(lab)) <= {0, -2}
        occ = a["multi"][f].max(0) > 0
        assert set(np.unique(a["multi"][f])) <= {0, 255}
        # a cell with occupancy in a layer >= 1 has z >= -1.875+... > -2 => positive single height
        assert (a["single"][f][a["multi"][f][2:].max(0) > 0] > 0).all()
        assert occ.sum() > 100


def _pack_records(batch, layout, rng):
    """Interleave the SoA batch into records of the given layout (padding filled with random bytes)."""
    n = len(batch["x"])
    rec = rng.integers(0, 256, size=(n, layout.stride), dtype=np.uint8)
    put = lambda off, a: rec.__setitem__((slice(None), slice(off, off + a.dtype.itemsize)),
                                         np.ascontiguousarray(a).view(np.uint8).reshape(n, a.dtype.itemsize))
    for name, off, t in (("x", layout.off_x, np.float32), ("y", layout.off_y, np.float32), ("z", layout.off_z, np.float32),
                         ("intensity", layout.off_intensity, np.float32), ("row", layout.off_row, np.uint16),
                         ("col", layout.off_col, np.uint16), ("label", layout.off_label, np.int16)):
        if off >= 0:
            put(off, np.asarray(batch[name], t))
    return rec.reshape(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["pcd26", "reordered_odd", "missing_fields"])
def test_packed_records_deinterleave_on_gpu(gens, synth, O, pkg, kind):
    """frames handed over as the interleaved records of a binary PCD payload (loadPCDFile, :730);
    the GPU de-interleave must give exactly what the SoA path gives (bit-exact vs the oracle)."""
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 7, first=300)           # 7 frames > host chunking of max_frames_per_batch=3
    rng = np.random.default_rng(5)
    if kind == "pcd26":
        lay = pkg.pcd_record_layout()
        assert (lay.stride, lay.off_x, lay.off_y, lay.off_z, lay.off_intensity, lay.off_row, lay.off_col, lay.off_label) == (26, 0, 4, 8, 12, 16, 18, 24)
    elif kind == "reordered_odd":                              # odd stride and offsets: byte-wise shared-memory reads
        lay = pkg.RecordLayout(37, 21, 3, 9, 29, 1, 33, 15)
    else:                                                      # intensity and label absent -> 0 (value-initialised)
        lay = pkg.RecordLayout(20, 0, 4, 8, -1, 12, 14, -1)
        batch = dict(batch); batch["intensity"] = np.zeros_like(batch["intensity"]); batch["label"] = np.zeros_like(batch["label"])
    ref = oracle_batch(O, sensor, batch)
    g = gens(sensor, max_frames_per_batch=3)
    rec = _pack_records(batch, lay, rng)
    out = g.process_packed_host(rec, batch["offsets"], lay)
    out["owner"] = pkg.owner_from_winner(out["winner"], batch["offsets"], np.asarray(batch["row"], np.uint16), np.asarray(batch["col"], np.uint16),
                                         g.params.horizon_scan, g.S)
    assert_same(out, ref, "packed " + kind)
    # ragged input: an empty frame in the middle and a one-point frame at the end
    fr = lambda f: {k: batch[k][batch["offsets"][f]:batch["offsets"][f + 1]] for k in FIELDS}
    empty = {k: batch[k][:0] for k in FIELDS}
    one = {k: batch[k][5:6] for k in FIELDS}
    rag = cat_frames([fr(0), empty, fr(1), one])
    out = g.process_packed_host(_pack_records(rag, lay, rng), rag["offsets"], lay)
    out["owner"] = pkg.owner_from_winner(out["winner"], rag["offsets"], np.asarray(rag["row"], np.uint16), np.asarray(rag["col"], np.uint16),
                                         g.params.horizon_scan, g.S)
    assert_same(out, oracle_batch(O, sensor, rag), "packed ragged " + kind)
    with pytest.raises(pkg.BevgenError, match="offset outside"):
        g.process_packed_host(rec, batch["offsets"], pkg.RecordLayout(lay.stride, lay.stride - 2, 4, 8, -1, 12, 14, -1))   # x (f32) overruns the record


def test_bird_view_map_of_batch_cloud_manip(gens, synth, O):
    """the optional `bvm` output = saveAsMat of BatchCloudManip.cpp:201-226 on the ground-removed ordered
    cloud (f32 max of z + 2 per 1 m cell, label == 0 skipped); bit-exact against the oracle, SoA and device chunking."""
    sensor = "HDL_64E"                                      # batch_cloud_manip hard-codes the HDL-64E shape (:12-13)
    batch = synth.make_batch(sensor, 5, first=20)
    sp = O.sensor(sensor)
    g = gens(sensor, max_frames_per_batch=2)
    offs = batch["offsets"]
    out = g.process_host(batch, out=g.alloc_outputs(5, n_total=int(offs[-1]), bvm=True))
    for f in range(5):
        fr = [batch[k][offs[f]:offs[f + 1]] for k in FIELDS]
        oc = O.order(sp, *fr)
        lab = O.mark_ground(sp, oc)[0]
        want = O.bvm(oc, lab)
        assert np.array_equal(out["label"][f], lab)
        got = out["bvm"][f]
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (f, int((got != want).sum()))
        assert want.max() > 0 and (want > 0).sum() > 200   # the map is not trivially empty


def test_projection_step_bit_exact(gens, O):
    """row / col of the range image as the extractors compute them, bit-exact against the oracle on
    random clouds plus zeros, axis points, huge / tiny / non-finite coordinates."""
    rng = np.random.default_rng(11)
    n = 300_001
    x = rng.normal(0, 30, n).astype(np.float32); y = rng.normal(0, 30, n).astype(np.float32); z = rng.normal(-1, 3, n).astype(np.float32)
    sp = [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e30, np.inf, -np.inf, np.nan, 1e-7, -1e-7]
    k = 0
    for a in sp:
        for b in sp:
            x[k], y[k], z[k] = a, b, sp[(k * 7) % len(sp)]; k += 1
    g = gens("OS1_64")
    got = g.project(0, x, y)
    row, col = O.project_mulran(x, y)
    assert np.array_equal(got["row"], row) and np.array_equal(got["col"], col)
    assert col.max() == 1024                                            # the col == Horizon_SCAN quirk is exercised
    got = g.project(1, x, y, z)
    nx, nz, row, col = O.project_oxford(x, y, z)
    for a, b in ((got["x"], nx), (got["z"], nz)):                      # negated in place; a NaN stays a NaN (payload / sign not compared)
        nan = np.isnan(b)
        assert np.array_equal(np.isnan(a), nan) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))
    assert np.array_equal(got["row"], row), int((got["row"] != row).sum())
    assert np.array_equal(got["col"], col), int((got["col"] != col).sum())
    assert col.max() < 1056 and set(np.unique(row)) == set(range(32))


@pytest.mark.parametrize("seed,kw", [(0, {}), (1, dict(start_negative=True)), (2, dict(n_rings=70, short_rings=(0, 1, 33))), (3, dict(n_rings=3, jitter=False))])
def test_kitti_ring_detection_bit_exact(gens, synth, O, seed, kw):
    """SURVEY 8(f)-2, KittiPointCloudSelect.cpp:188-243: row / col per point of a raw scan, equal to the oracle - short rings
    whose crossing is ignored, spurious sign flips, a scan that starts below 0 degrees, more than 64 rings."""
    x, y, _ = synth.make_kitti_scan(seed, **kw)
    row, col = O.project_kitti(x, y)
    got = gens("HDL_64E").project(2, x, y)
    assert np.array_equal(got["row"], row), int((got["row"] != row).sum())
    assert np.array_equal(got["col"], col), int((got["col"] != col).sum())
    placed = row != 0xFFFF
    assert row[0] == 0xFFFF and placed.sum() > 1000 and col[placed].max() < 2083 and row[placed].max() < 64
    if kw.get("n_rings", 64) == 64 and not kw.get("start_negative"):
        assert row[placed].max() == 61          # 64 rings, two short ones merged into their successors


def test_top_flatten_segmented_selection(gens, synth, O):
    """extractTopAndFlatten on the GPU (stable radix sort by cell / height + per-cell quota) against the
    oracle: a real ground-removed frame, a wide random cloud with many equal heights (ties keep input order), tiny inputs."""
    g = gens("HDL_64E")
    sp = O.sensor("HDL_64E")
    f = synth.make_frame("HDL_64E", 77)
    oc = O.order(sp, *[f[k] for k in FIELDS])
    lab = O.mark_ground(sp, oc)[0]                                   # what non_ground_point_cloud/*.pcd holds
    cases = [(oc["x"], oc["y"], oc["z"], lab)]
    rng = np.random.default_rng(21)
    n = 300_007
    cases.append((rng.uniform(-130, 130, n).astype(np.float32), rng.uniform(-130, 130, n).astype(np.float32),
                  (rng.integers(-8, 40, n) * 0.25).astype(np.float32), rng.integers(-2, 3, n).astype(np.int16)))   # quantised heights: ties
    cases.append((np.zeros(25, np.float32), np.zeros(25, np.float32), np.r_[np.arange(24, dtype=np.float32), -0.0].astype(np.float32), np.ones(25, np.int16)))
    cases.append((np.zeros(0, np.float32),) * 3 + (np.zeros(0, np.int16),))
    for x, y, z, l in cases:
        wx, wy, wi = O.top_flatten(x, y, z, l)
        gx, gy, gi = g.top_flatten(x, y, z, l)
        assert len(gi) == len(wi), (len(gi), len(wi))
        assert np.array_equal(gi, wi), int((gi != wi).sum())
        assert np.array_equal(gx.view(np.uint32), wx.view(np.uint32)) and np.array_equal(gy.view(np.uint32), wy.view(np.uint32))
    assert len(O.top_flatten(*cases[0])[2]) > 1000 and len(O.top_flatten(*cases[1])[2]) > 10000


# ---- round 2: compact staging format, libm overload switch, diagnostics, leak / ring fixes -------------------------
def _compact(pkg, g, batch):
    return dict(x=batch["x"], y=batch["y"], z=batch["z"], offsets=batch["offsets"],
                meta=pkg.pack_meta(g.params, batch["row"], batch["col"], batch["intensity"], batch["label"]))


@pytest.mark.parametrize("sensor", ["HDL_32E", "OS1_64", "HDL_64E"])
def test_compact_host_path_bit_exact_after_expand(gens, pkg, synth, O, sensor):
    """bevgen_process_host_compact: 16 B/point in, ground bits + bit planes out; after the host-side expansion (a change of
    representation only) every output equals the oracle's - synthetic frames (chunked: 7 > 3) and the random frames."""
    sp = O.sensor(sensor)
    g = gens(sensor, max_frames_per_batch=3, max_points_per_frame=sp.S * 2)
    for batch in (synth.make_batch(sensor, 7, first=400), cat_frames(cases.random_unstructured_frames(sp) + [cases.hot_cell_frame(sp)])):
        cout = g.process_host_compact(_compact(pkg, g, batch))
        assert cout["planes"].shape[1:] == (3, 224, 224) and cout["ground"].shape[1] == (sp.S + 31) // 32
        assert_same(g.compact_to_reference_layout(cout, batch), oracle_batch(O, sensor, batch), "compact " + sensor)
    # pinned buffers (the fully asynchronous route) give the same bytes
    batch = synth.make_batch(sensor, 4, first=410)
    cb = _compact(pkg, g, batch)
    pin = {k: pkg.pinned_empty(np.asarray(cb[k]).shape, np.asarray(cb[k]).dtype) for k in ("x", "y", "z", "meta")}
    for k in pin:
        pin[k][...] = cb[k]
    pin["offsets"] = cb["offsets"]
    out = g.alloc_outputs_compact(4, pinned=True, n_total=int(batch["offsets"][-1]))
    g.process_host_compact(pin, out=out)
    assert_same(g.compact_to_reference_layout(out, batch), oracle_batch(O, sensor, batch), "compact pinned " + sensor)
    for a in list(pin.values())[:4] + list(out.values()):
        pkg.pinned_free(a)
    # write-combined input staging (bevgen_host_alloc_wc): written once by the host, only ever read by the copy engine
    wc = {k: pkg.pinned_empty(np.asarray(cb[k]).shape, np.asarray(cb[k]).dtype, write_combined=True) for k in ("x", "y", "z", "meta")}
    for k in wc:
        wc[k][...] = cb[k]
    wc["offsets"] = cb["offsets"]
    assert_same(g.compact_to_reference_layout(g.process_host_compact(wc), batch), oracle_batch(O, sensor, batch), "compact write-combined " + sensor)
    for a in list(wc.values())[:4]:
        pkg.pinned_free(a)


def test_libm_double_switch_and_diag(pkg, O):
    """bevgen_set_libm(1): the C double atan2 / sqrt overload set (the reference built without <math.h> in its include
    tree); bevgen_set_diag counts the borderline pairs and the pairs on which the two overload sets disagree."""
    sensor = "HDL_32E"
    sp = O.sensor(sensor)
    batch = cat_frames(cases.borderline_frames(sp))
    ref_f = oracle_batch(O, sensor, batch); ref_d = oracle_batch(O, sensor, batch, double_libm=True)
    assert (ref_f["label"] != ref_d["label"]).any()
    g = pkg.BevGen(sensor, device=0, max_frames_per_batch=2, max_points_per_frame=sp.S * 2)
    try:
        g.set_diag(True)
        assert_same(g.process_host(batch), ref_f, "float libm + diag")
        d = g.get_diag()
        assert d["borderline_pairs"] > 1000 and d["float_double_disagree"] > 0, d
        g.set_libm(True)
        assert_same(g.process_host(batch), ref_d, "double libm")
        g.set_diag(False)
        assert_same(g.process_host(batch), ref_d, "double libm, diag off")
        assert g.get_diag() == dict(borderline_pairs=0, float_double_disagree=0)
        g.set_libm(False)
        assert_same(g.process_host(batch), ref_f, "back to float")
    finally:
        g.close()


def test_packed_records_stride_256(gens, synth, O, pkg):
    """records of 192..256 bytes need more than 48 KB of dynamic shared memory in k_unpack_records."""
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 3, first=310)
    g = gens(sensor, max_frames_per_batch=3)
    rng = np.random.default_rng(6)
    for lay in (pkg.RecordLayout(256, 100, 8, 248, 200, 2, 252, 0), pkg.RecordLayout(193, 1, 5, 9, 13, 17, 19, 21)):
        out = g.process_packed_host(_pack_records(batch, lay, rng), batch["offsets"], lay)
        out["owner"] = pkg.owner_from_winner(out["winner"], batch["offsets"], np.asarray(batch["row"], np.uint16), np.asarray(batch["col"], np.uint16),
                                             g.params.horizon_scan, g.S)
        assert_same(out, oracle_batch(O, sensor, batch), "stride %d" % lay.stride)


def test_create_failure_releases_everything(pkg):
    """a bevgen_create that fails part-way (device memory exhausted) must free what it had allocated."""
    import torch
    torch.cuda.init()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(40):
        with pytest.raises(pkg.BevgenError):
            pkg.BevGen("HDL_64E", device=0, max_frames_per_batch=65535, max_points_per_frame=4_000_000)
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 64 << 20, (free0, free1)
    g = pkg.BevGen("HDL_32E", device=0, max_frames_per_batch=2)      # the device is still usable
    g.close()


def test_ring_holds_max_frames_per_batch(pkg, synth, O):
    """the submit / collect ring holds max_frames_per_batch frames (it used to stop at 8)."""
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 11, first=60)
    ref = oracle_batch(O, sensor, batch)
    offs = batch["offsets"]
    fr = lambda f: {k: batch[k][offs[f]:offs[f + 1]] for k in FIELDS}
    g = pkg.BevGen(sensor, device=0, max_frames_per_batch=11)
    try:
        for f in range(11):
            g.submit(f, fr(f))
        with pytest.raises(pkg.BevgenError, match="ring full"):
            g.submit(11, fr(0))
        for f in (10, 0, 5, 1, 2, 3, 4, 6, 7, 8, 9):
            o = g.collect(f, fr(f))
            assert_same({k: v[None] for k, v in o.items() if k != "winner"}, {k: v[f:f + 1] for k, v in ref.items()}, "ring %d" % f)
    finally:
        g.close()


def test_cloud_manip_2m_points_device_resident(gens, O):
    """BASELINE config #5 at its stated size: 2 M points, 60 % in a 3 m blob around the origin; the device-resident entry
    point, the independent optional outputs and the empty cloud."""
    import torch
    g = gens("HDL_32E", max_frames_per_batch=4, max_points_per_frame=33792 * 2)
    rng = np.random.default_rng(55)
    n = 2_000_000
    hot = rng.random(n) < 0.6
    x = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    y = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    z = rng.uniform(-2, 10, n).astype(np.float32)
    th = np.float32(np.float32(37.0) / np.float32(180.0) * np.pi)
    c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
    otx, oty, otz = O.transform(rt, x, y, z)
    want_i = O.save_as_mat(x, y, z); want_o = O.save_as_mat(otx, oty, otz)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(v).to(dev) for k, v in (("x", x), ("y", y), ("z", z))}
    for k in ("tx", "ty", "tz"):
        d[k] = torch.empty(n, dtype=torch.float32, device=dev)
    d["bev_in"] = torch.full((201, 201), 7.0, dtype=torch.float32, device=dev); d["bev_out"] = torch.full((201, 201), 7.0, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    g.cloud_manip_device(n, rt, {k: v.data_ptr() for k, v in d.items()})
    g.sync()
    for a, b in (("tx", otx), ("ty", oty), ("tz", otz)):
        assert np.array_equal(d[a].cpu().numpy().view(np.uint32), b.view(np.uint32)), a
    assert np.array_equal(d["bev_in"].cpu().numpy(), want_i) and np.array_equal(d["bev_out"].cpu().numpy(), want_o)
    # host form: each output optional on its own; an empty cloud (NULL arrays) gives zero grids
    import ctypes as C
    L = pkg_lib = __import__("pcpt_b200").lib()
    bo = np.empty((201, 201), np.float32); ty = np.empty(n, np.float32)
    rc = L.bevgen_cloud_manip(g._ctx, C.c_int64(n), rt.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                              z.ctypes.data_as(C.c_void_p), None, ty.ctypes.data_as(C.c_void_p), None, None, bo.ctypes.data_as(C.c_void_p))
    assert rc == 0 and np.array_equal(ty.view(np.uint32), oty.view(np.uint32)) and np.array_equal(bo, want_o)
    bi = np.full((201, 201), 3.0, np.float32)
    rc = L.bevgen_cloud_manip(g._ctx, C.c_int64(0), rt.ctypes.data_as(C.c_void_p), None, None, None, None, None, None, bi.ctypes.data_as(C.c_void_p), None)
    assert rc == 0 and not bi.any()


def test_null_arrays_are_refused_not_launched(pkg, synth):
    """A NULL point / output array must come back as an error from the C-ABI instead of reaching a kernel (only
    bevgen_outputs.bvm is optional); an empty frame may be submitted with NULL arrays; the context stays usable."""
    import ctypes as C
    sensor = "HDL_32E"
    batch = synth.make_batch(sensor, 2, first=3)
    g = pkg.BevGen(sensor, device=0, max_frames_per_batch=2)
    L = pkg.lib()
    try:
        offs = np.ascontiguousarray(batch["offsets"], np.int64)
        out = g.alloc_outputs(2, n_total=int(offs[-1]))
        arrs = [np.ascontiguousarray(batch[k]) for k in FIELDS]
        ptrs = [C.c_void_p(a.ctypes.data) for a in arrs]
        full = pkg.Outputs(*[C.c_void_p(out[k].ctypes.data) for k in ("label", "winner", "single", "multi")], None)
        for hole in range(7):       # one input array missing
            pts = pkg.Points(*[None if i == hole else p for i, p in enumerate(ptrs)])
            assert L.bevgen_process_host(g._ctx, C.c_int(2), C.c_void_p(offs.ctypes.data), C.byref(pts), C.byref(full)) == -1
            assert b"null array" in L.bevgen_last_error()
        pts = pkg.Points(*ptrs)
        for hole in range(4):       # one output array missing
            o = pkg.Outputs(*[None if i == hole else C.c_void_p(out[k].ctypes.data) for i, k in enumerate(("label", "winner", "single", "multi"))], None)
            assert L.bevgen_process_host(g._ctx, C.c_int(2), C.c_void_p(offs.ctypes.data), C.byref(pts), C.byref(o)) == -1
            assert L.bevgen_process_device(g._ctx, C.c_int(2), C.c_void_p(offs.ctypes.data), C.byref(pts), C.byref(o)) == -1
        assert L.bevgen_submit(g._ctx, C.c_int(0), C.c_int(5), None, None, None, None, None, None, None) == -1
        assert L.bevgen_submit(g._ctx, C.c_int(0), C.c_int(0), None, None, None, None, None, None, None) == 0   # empty frame
        lab = np.empty(g.S, np.int16); single = np.empty((224, 224), np.uint8)
        assert L.bevgen_collect(g._ctx, C.c_int(0), C.c_void_p(lab.ctypes.data), None, C.c_void_p(single.ctypes.data), None) == 0
        assert not lab.any() and not single.any()
        good = g.process_host(batch)                                    # and the context still computes
        assert good["single"].any()
    finally:
        g.close()
