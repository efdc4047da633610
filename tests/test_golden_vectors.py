"""Golden vectors generated from the reference's own source (tests/golden/make_bev_golden.py -> bev_golden.json,
bev_golden_small.npz): the CPU test holds the oracle to them (no /root/reference needed), the GPU test holds the CUDA
path to them — so on the GPU box the product is compared with reference-generated outputs, not only with our oracle."""
import json
import os

import numpy as np
import pytest

import cases
from conftest import FIELDS, cat_frames

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "bev_golden.json")))
SMALL = np.load(os.path.join(HERE, "golden", "bev_golden_small.npz"))


def _check(cid, got, what):
    g = META["frames"][cid]
    for k in ("owner", "label", "single", "multi"):
        if cases.digest(got[k]) != g[k]:
            key = cid + ":" + k
            where = ""
            if key in SMALL.files:
                bad = np.argwhere(np.asarray(got[k]).reshape(SMALL[key].shape) != SMALL[key])
                where = " (%d mismatches, first at %s)" % (len(bad), bad[:1].tolist())
            raise AssertionError("%s: %s of %s differs from the reference-generated golden vector%s" % (what, k, cid, where))


def test_oracle_matches_reference_generated_vectors(O, synth):
    frames = cases.golden_frame_list(synth, O)
    assert sorted(c for c, _, _ in frames) == sorted(META["frames"])
    for cid, sensor, f in frames:
        o = O.frame(O.sensor(sensor), *[f[k] for k in FIELDS])
        _check(cid, o, "oracle")
        od = O.frame(O.sensor(sensor), *[f[k] for k in FIELDS], double_libm=True)
        assert cases.digest(od["label"]) == META["frames"][cid]["label_double_libm"], cid
    for K, seed, step in cases.GOLDEN_LABEL_SETS:
        g = META["labels"]["K%d_s%d" % (K, seed)]
        xyz = synth.make_poses(K, seed=seed, step=step)
        mi, _ = O.select_major(xyz)
        lab, _, _ = O.labels(xyz, mi)
        assert len(mi) == g["M"] and cases.digest(mi) == g["major"] and cases.digest(lab) == g["labels"], (K, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("sensor", ["HDL_32E", "OS1_64", "HDL_64E"])
def test_cuda_matches_reference_generated_vectors(pkg, synth, O, sensor):
    frames = [(c, f) for c, s, f in cases.golden_frame_list(synth, O) if s == sensor]
    sp = O.sensor(sensor)
    g = pkg.BevGen(sensor, device=0, max_frames_per_batch=4, max_points_per_frame=sp.S * 2)
    try:
        out = g.process_host(cat_frames([f for _, f in frames]))
        for i, (cid, _) in enumerate(frames):
            _check(cid, {k: out[k][i] for k in ("owner", "label", "single", "multi")}, "CUDA")
    finally:
        g.close()


@pytest.mark.gpu
def test_cuda_labels_match_reference_generated_vectors(pkg, synth):
    """K = 10 000 keyframes / M = 457 majors is BASELINE configs[2],[3]'s label stage at full size."""
    g = pkg.BevGen("HDL_32E", device=0, max_frames_per_batch=2)
    try:
        for K, seed, step in cases.GOLDEN_LABEL_SETS:
            gold = META["labels"]["K%d_s%d" % (K, seed)]
            xyz = synth.make_poses(K, seed=seed, step=step)
            mi, _ = g.select_major(xyz)
            assert len(mi) == gold["M"] and cases.digest(mi) == gold["major"], (K, seed)
            lab, _, _ = g.labels(xyz, mi)
            assert cases.digest(lab) == gold["labels"], (K, seed)
            h = K // 3                                                  # row split as the multi-GPU label stage does it
            parts = [g.labels(xyz, mi, a, b)[0] for a, b in ((0, h), (h, 2 * h), (2 * h, K))]
            assert cases.digest(np.concatenate(parts)) == gold["labels"], (K, seed)
    finally:
        g.close()


# ---- extractTopAndFlatten: vectors generated from TopPartRegistration.cpp itself -------------------------
TOP = np.load(os.path.join(HERE, "golden", "top_flatten_golden.npz"))


def _check_top(name, got_x, got_y, what):
    key = name.replace(" ", "_")
    gx, gy = np.asarray(got_x, np.float32), np.asarray(got_y, np.float32)
    assert len(gx) == int(TOP[key + ":n"]), (what, name, len(gx), int(TOP[key + ":n"]))
    if key + ":x" in TOP.files:
        assert np.array_equal(gx.view(np.uint32), TOP[key + ":x"].view(np.uint32)), (what, name, "x")
        assert np.array_equal(gy.view(np.uint32), TOP[key + ":y"].view(np.uint32)), (what, name, "y")
    assert cases.digest(np.concatenate([gx, gy])) == str(TOP[key + ":sha256"]), (what, name)


def test_oracle_top_flatten_matches_reference_generated_vectors(O, synth):
    names = []
    for name, x, y, z, lab in cases.top_flatten_cases(O, synth):
        ox, oy, _ = O.top_flatten(x, y, z, lab)
        _check_top(name, ox, oy, "oracle")
        names.append(name.replace(" ", "_"))
    assert sorted(k[:-2] for k in TOP.files if k.endswith(":n")) == sorted(names)


@pytest.mark.gpu
def test_cuda_top_flatten_matches_reference_generated_vectors(pkg, synth, O):
    g = pkg.BevGen("HDL_64E", device=0, max_frames_per_batch=2)
    try:
        for name, x, y, z, lab in cases.top_flatten_cases(O, synth):
            gx, gy, gi = g.top_flatten(x, y, z, lab)
            _check_top(name, gx, gy, "CUDA")
            assert np.array_equal(x[gi].view(np.uint32), np.asarray(gx, np.float32).view(np.uint32)), name
    finally:
        g.close()


# ---- the extractors' projection step: vectors generated from {Mulran,Oxford,Kitti}PointCloudSelect.cpp themselves --
PROJ = json.load(open(os.path.join(HERE, "golden", "projection_golden.json")))


def _check_projection(project, kitti, what, sfx=""):
    """project(kind, x, y, z) -> dict(row, col, x, z), the C-ABI's shape; kitti(seed) -> the scan's x, y, z."""
    x, y, z = cases.projection_cloud()
    g = PROJ["mulran" + sfx]; n = g["n"]
    r = project(0, x[:n], y[:n], None)
    assert cases.digest(r["row"]) == g["row"] and cases.digest(r["col"]) == g["col"], (what, "mulran")
    g = PROJ["oxford" + sfx]
    assert g["n"] == len(x)
    r = project(1, x, y, z)
    for k in ("row", "col", "x", "z"):
        assert cases.digest(r[k]) == g[k], (what, "oxford", k)
    for seed, _ in cases.KITTI_SCANS:
        g = PROJ["kitti_%d%s" % (seed, sfx)]
        kx, ky, kz = kitti(seed)
        assert len(kx) == g["n"]
        r = project(2, kx, ky, None)
        s = cases.kitti_structured(kx, ky, kz, r["row"], r["col"])
        assert int((s["label"] == -2).sum()) == g["written_slots"], (what, "kitti", seed)
        assert cases.structured_digest(s) == g["structured"], (what, "kitti", seed)


def _kitti_scans(synth):
    kws = dict(cases.KITTI_SCANS)
    return lambda seed: synth.make_kitti_scan(seed, **kws[seed])


@pytest.mark.parametrize("double_libm", [False, True])
def test_oracle_projection_matches_reference_generated_vectors(O, synth, double_libm):
    def project(kind, x, y, z):
        if kind == 0:
            row, col = O.project_mulran(x, y, double_libm=double_libm)
            return dict(row=row, col=col)
        if kind == 1:
            nx, nz, row, col = O.project_oxford(x, y, z, double_libm=double_libm)
            return dict(row=row, col=col, x=nx, z=nz)
        row, col = O.project_kitti(x, y, double_libm=double_libm)
        return dict(row=row, col=col)
    _check_projection(project, _kitti_scans(synth), "oracle", "_double_libm" if double_libm else "")


@pytest.mark.gpu
def test_cuda_projection_matches_reference_generated_vectors(pkg, synth):
    g = pkg.BevGen("HDL_64E", device=0, max_frames_per_batch=2)
    try:
        _check_projection(lambda kind, x, y, z: g.project(kind, x, y, z), _kitti_scans(synth), "CUDA")
    finally:
        g.close()
