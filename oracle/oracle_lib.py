"""ctypes wrapper around the CPU oracle (oracle/bevgen_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (point-cloud-preprocessing-tools_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GRID = 224
LAYERS = 24
SECT_R, SECT_C = 75, 50


class Sensor(C.Structure):
    _fields_ = [("n_scan", C.c_int32), ("horizon_scan", C.c_int32), ("ground_upper_scan", C.c_int32),
                ("height_res", C.c_float)]

    @property
    def S(self):
        return self.n_scan * self.horizon_scan


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    so = os.path.join(_HERE, "libbevgen_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "bevgen_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def lib(double_libm=False):
    key = "dbl" if double_libm else "flt"
    if key not in _libs:
        build()
        name = "libbevgen_oracle_dbl.so" if double_libm else "libbevgen_oracle.so"
        _libs[key] = C.CDLL(os.path.join(_HERE, name))
        _libs[key].oracle_select_major.restype = C.c_int
        _libs[key].oracle_sensor_params.restype = C.c_int
        _libs[key].oracle_atan2f.restype = C.c_float
        _libs[key].oracle_atan2f.argtypes = [C.c_float, C.c_float]
        _libs[key].oracle_angle_deg.restype = C.c_float
        _libs[key].oracle_angle_deg.argtypes = [C.c_float, C.c_float, C.c_float]
    return _libs[key]


def ref_lib():
    """oracle/_ref/libnanoflann_ref.so — the reference's own vendored KD-tree, or None if it was never built."""
    p = os.path.join(_HERE, "_ref", "libnanoflann_ref.so")
    if not os.path.exists(p):
        return None
    if "ref" not in _libs:
        _libs["ref"] = C.CDLL(p)
        _libs["ref"].ref_knn.restype = C.c_int
    return _libs["ref"]


def sensor(name):
    s = Sensor()
    rc = lib().oracle_sensor_params(name.encode(), C.byref(s))
    if rc < 0:
        raise ValueError("Unknown sensor type: %s!" % name)
    return s


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, _p(a, C.c_float)


def order(sp, x, y, z, intensity, row, col, label):
    S = sp.S
    x, px = _f(x); y, py = _f(y); z, pz = _f(z); it, pi = _f(intensity)
    row = np.ascontiguousarray(row, np.uint16); col = np.ascontiguousarray(col, np.uint16)
    label = np.ascontiguousarray(label, np.int16)
    out = {k: np.empty(S, np.float32) for k in ("x", "y", "z", "intensity")}
    out["label"] = np.empty(S, np.int16); out["owner"] = np.empty(S, np.uint32)
    lib().oracle_order(C.byref(sp), C.c_int64(len(x)), px, py, pz, pi, _p(row, C.c_uint16), _p(col, C.c_uint16),
                       _p(label, C.c_int16), _p(out["x"], C.c_float), _p(out["y"], C.c_float), _p(out["z"], C.c_float),
                       _p(out["intensity"], C.c_float), _p(out["label"], C.c_int16), _p(out["owner"], C.c_uint32))
    return out


def mark_ground(sp, oc, double_libm=False):
    """oc = dict from order(); returns (label_out, gm_after_loop1, gm_final, avg[75,50]); oc is not modified."""
    S = sp.S
    lab = oc["label"].copy()
    gm = np.empty(S, np.int8); gmf = np.empty(S, np.int8); avg = np.empty(SECT_R * SECT_C, np.float32)
    lib(double_libm).oracle_mark_ground(C.byref(sp), _p(oc["x"], C.c_float), _p(oc["y"], C.c_float), _p(oc["z"], C.c_float),
                                        _p(oc["intensity"], C.c_float), _p(lab, C.c_int16), _p(gm, C.c_int8),
                                        _p(gmf, C.c_int8), _p(avg, C.c_float))
    return lab, gm.reshape(sp.n_scan, sp.horizon_scan), gmf.reshape(sp.n_scan, sp.horizon_scan), avg.reshape(SECT_R, SECT_C)


def multi_bev(sp, oc, label):
    m = np.empty(LAYERS * GRID * GRID, np.uint8)
    label = np.ascontiguousarray(label, np.int16)
    lib().oracle_multi_bev(C.byref(sp), _p(oc["x"], C.c_float), _p(oc["y"], C.c_float), _p(oc["z"], C.c_float),
                           _p(label, C.c_int16), _p(m, C.c_uint8))
    return m.reshape(LAYERS, GRID, GRID)


def single_bev(sp, oc, label):
    m = np.empty(GRID * GRID, np.uint8)
    label = np.ascontiguousarray(label, np.int16)
    lib().oracle_single_bev(C.byref(sp), _p(oc["x"], C.c_float), _p(oc["y"], C.c_float), _p(oc["z"], C.c_float),
                            _p(label, C.c_int16), _p(m, C.c_uint8))
    return m.reshape(GRID, GRID)


def frames(sp, offsets, x, y, z, intensity, row, col, label, n_threads=1, double_libm=False):
    """Batch of frames (concatenated SoA + offsets[F+1]) -> dict(label, owner, single, multi)."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    F = len(offsets) - 1
    S = sp.S
    x, px = _f(x); y, py = _f(y); z, pz = _f(z); it, pi = _f(intensity)
    row = np.ascontiguousarray(row, np.uint16); col = np.ascontiguousarray(col, np.uint16)
    label = np.ascontiguousarray(label, np.int16)
    out = dict(label=np.empty((F, S), np.int16), owner=np.empty((F, S), np.uint32),
               single=np.empty((F, GRID, GRID), np.uint8), multi=np.empty((F, LAYERS, GRID, GRID), np.uint8))
    lib(double_libm).oracle_frames(C.byref(sp), C.c_int(F), _p(offsets, C.c_int64), px, py, pz, pi, _p(row, C.c_uint16),
                                   _p(col, C.c_uint16), _p(label, C.c_int16), _p(out["label"], C.c_int16),
                                   _p(out["owner"], C.c_uint32), _p(out["single"], C.c_uint8), _p(out["multi"], C.c_uint8),
                                   C.c_int(n_threads))
    return out


def frame(sp, x, y, z, intensity, row, col, label, **kw):
    o = frames(sp, [0, len(x)], x, y, z, intensity, row, col, label, **kw)
    return {k: v[0] for k, v in o.items()}


def select_major(xyz):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    K = len(xyz)
    mi = np.empty(max(K, 1), np.int32); ov = np.empty(max(K, 1), np.int32)
    M = lib().oracle_select_major(C.c_int(K), _p(xyz, C.c_float), _p(mi, C.c_int32), _p(ov, C.c_int32))
    return mi[:M].copy(), ov[:K].copy()


def labels(xyz, major_idx):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    major_idx = np.ascontiguousarray(major_idx, np.int32)
    K, M = len(xyz), len(major_idx)
    lab = np.empty((K, M), np.float32); nn = np.empty((K, 2), np.int32); w = np.empty((K, 2), np.float32)
    lib().oracle_labels(C.c_int(K), _p(xyz, C.c_float), C.c_int(M), _p(major_idx, C.c_int32), _p(lab, C.c_float),
                        _p(nn, C.c_int32), _p(w, C.c_float))
    return lab, nn, w


def transform(rt, x, y, z):
    rt = np.ascontiguousarray(rt, np.float32).reshape(12)
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    o = [np.empty(len(x), np.float32) for _ in range(3)]
    lib().oracle_transform(C.c_int64(len(x)), _p(rt, C.c_float), px, py, pz, *[_p(a, C.c_float) for a in o])
    return o


def save_as_mat(x, y, z):
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    m = np.empty(201 * 201, np.float32)
    lib().oracle_save_as_mat(C.c_int64(len(x)), px, py, pz, _p(m, C.c_float))
    return m.reshape(201, 201)


def bvm(oc, label):
    """batch_cloud_manip's bird-view map (BatchCloudManip.cpp:201-226) of the ordered cloud oc with post-ground labels."""
    x, px = _f(oc["x"]); y, py = _f(oc["y"]); z, pz = _f(oc["z"])
    label = np.ascontiguousarray(label, np.int16)
    m = np.empty(201 * 201, np.float32)
    lib().oracle_bvm(C.c_int64(len(x)), px, py, pz, _p(label, C.c_int16), _p(m, C.c_float))
    return m.reshape(201, 201)


def project_mulran(x, y, double_libm=False):
    """MulranPointCloudSelect.cpp:112-126 -> (row, col)."""
    x, px = _f(x); y, py = _f(y)
    row = np.empty(len(x), np.uint16); col = np.empty(len(x), np.uint16)
    lib(double_libm).oracle_project_mulran(C.c_int64(len(x)), px, py, _p(row, C.c_uint16), _p(col, C.c_uint16))
    return row, col


def project_kitti(x, y, double_libm=False):
    """KittiPointCloudSelect.cpp:188-243 -> (row, col); 0xFFFF = not placed."""
    x, px = _f(x); y, py = _f(y)
    row = np.empty(len(x), np.uint16); col = np.empty(len(x), np.uint16)
    lib(double_libm).oracle_project_kitti(C.c_int64(len(x)), px, py, _p(row, C.c_uint16), _p(col, C.c_uint16))
    return row, col


def project_oxford(x, y, z, double_libm=False):
    """OxfordPointCloudSelect.cpp:201-219 -> (x_negated, z_negated, row, col)."""
    x = np.array(x, np.float32); z = np.array(z, np.float32); y, py = _f(y)
    row = np.empty(len(x), np.uint16); col = np.empty(len(x), np.uint16)
    lib(double_libm).oracle_project_oxford(C.c_int64(len(x)), _p(x, C.c_float), py, _p(z, C.c_float), _p(row, C.c_uint16), _p(col, C.c_uint16))
    return x, z, row, col


def top_flatten(x, y, z, label):
    """extractTopAndFlatten (TopPartRegistration.cpp:79-141) -> (out_x, out_y, source_index)."""
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    label = np.ascontiguousarray(label, np.int16)
    n = len(x)
    ox = np.empty(max(n, 1), np.float32); oy = np.empty(max(n, 1), np.float32); oi = np.empty(max(n, 1), np.uint32)
    f = lib().oracle_top_flatten
    f.restype = C.c_int64
    m = f(C.c_int64(n), px, py, pz, _p(label, C.c_int16), _p(ox, C.c_float), _p(oy, C.c_float), _p(oi, C.c_uint32))
    return ox[:m].copy(), oy[:m].copy(), oi[:m].copy()


# ---- reference KD-tree (oracle/_ref) ---------------------------------------------------------------
def ref_knn_many(pts, qs, k):
    r = ref_lib()
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3); qs = np.ascontiguousarray(qs, np.float32).reshape(-1, 3)
    idx = np.zeros((len(qs), k), np.uint64); d = np.zeros((len(qs), k), np.float32)
    r.ref_knn_many(_p(pts, C.c_float), C.c_int(len(pts)), _p(qs, C.c_float), C.c_int(len(qs)), C.c_int(k),
                   _p(idx, C.c_uint64), _p(d, C.c_float))
    return idx.astype(np.int64), d


def ref_knn(pts, q, k):
    r = ref_lib()
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3); q = np.ascontiguousarray(q, np.float32).reshape(3)
    idx = np.zeros(k, np.uint64); d = np.zeros(k, np.float32)
    r.ref_knn(_p(pts, C.c_float), C.c_int(len(pts)), _p(q, C.c_float), C.c_int(k), _p(idx, C.c_uint64), _p(d, C.c_float))
    return idx.astype(np.int64), d


# ---- the reference's own translation unit, compiled against oracle/stub (oracle/_ref/libbevgen_ref*.so) -----------
def ref_bevgen_lib(double_libm=False):
    """oracle/_ref/libbevgen_ref.so (or _dbl: built with -DSTUB_NO_MATH_H so atan2/sqrt bind to the C double functions),
    or None when it was never built (it is built here, where /root/reference exists, and travels to the GPU box)."""
    key = "refbev_dbl" if double_libm else "refbev"
    if key not in _libs:
        p = os.path.join(_HERE, "_ref", "libbevgen_ref_dbl.so" if double_libm else "libbevgen_ref.so")
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.ref_get_distance.restype = C.c_float
        _libs[key] = L
    return _libs[key]


def ref_math_overloads(double_libm=False):
    L = ref_bevgen_lib(double_libm)
    a = [C.c_int() for _ in range(4)]
    L.ref_math_overloads(*[C.byref(v) for v in a])
    return dict(zip(("atan2", "sqrt", "abs", "round"), ["double" if v.value == 1 else "float" if v.value == 0 else "int" for v in a]))


def ref_frame(sensor_name, x, y, z, intensity, row, col, label, t=None, double_libm=False, want_csv=False):
    """One frame through getOrderedCloud -> markGroundPoints -> computeAndSaveMultiBev -> computeAndSaveSingleBev as the
    reference's own source computes them.  Returns dict(x,y,z,intensity,row,col,t,label [S] = the ordered cloud,
    ground_mat [N,H] i8, single, multi, bin_equal, csv)."""
    L = ref_bevgen_lib(double_libm)
    sp4 = (C.c_int32 * 4)()
    if L.ref_set_sensor(sensor_name.encode(), sp4) != 0:
        raise ValueError("Unknown sensor type: %s!" % sensor_name)
    N, H = sp4[0], sp4[1]
    S = N * H
    x, px = _f(x); y, py = _f(y); z, pz = _f(z); it, pi = _f(intensity)
    row = np.ascontiguousarray(row, np.uint16); col = np.ascontiguousarray(col, np.uint16)
    label = np.ascontiguousarray(label, np.int16)
    tt = None if t is None else np.ascontiguousarray(t, np.uint32)
    o = dict(x=np.empty(S, np.float32), y=np.empty(S, np.float32), z=np.empty(S, np.float32), intensity=np.empty(S, np.float32),
             row=np.empty(S, np.uint16), col=np.empty(S, np.uint16), t=np.empty(S, np.uint32), label=np.empty(S, np.int16),
             ground_mat=np.empty(S, np.int8), single=np.empty(GRID * GRID, np.uint8), multi=np.empty(LAYERS * GRID * GRID, np.uint8))
    beq = C.c_int(0); clen = C.c_int64(0)
    cap = 400000 if want_csv else 0
    cbuf = C.create_string_buffer(cap) if want_csv else None
    rc = L.ref_frame(C.c_int64(len(x)), px, py, pz, pi, _p(row, C.c_uint16), _p(col, C.c_uint16),
                     _p(tt, C.c_uint32) if tt is not None else None, _p(label, C.c_int16),
                     _p(o["x"], C.c_float), _p(o["y"], C.c_float), _p(o["z"], C.c_float), _p(o["intensity"], C.c_float),
                     _p(o["row"], C.c_uint16), _p(o["col"], C.c_uint16), _p(o["t"], C.c_uint32), _p(o["label"], C.c_int16),
                     _p(o["ground_mat"], C.c_int8), _p(o["single"], C.c_uint8), _p(o["multi"], C.c_uint8),
                     C.byref(beq), cbuf, C.c_int64(cap), C.byref(clen))
    if rc != S:
        raise RuntimeError("ref_frame returned %d (expected S=%d)" % (rc, S))
    o["ground_mat"] = o["ground_mat"].reshape(N, H)
    o["single"] = o["single"].reshape(GRID, GRID); o["multi"] = o["multi"].reshape(LAYERS, GRID, GRID)
    o["bin_equal"] = bool(beq.value)
    o["csv"] = cbuf.raw[:clen.value] if want_csv and clen.value > 0 else None
    return o


def ref_select_and_label(xyz, want_labels=True):
    """selectMajorFrames + getKeyFrameLabel of the reference's own source -> (major_idx, labels [K,M] or None).
    (The reference prints its progress lines to stdout.)"""
    L = ref_bevgen_lib()
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    K = len(xyz)
    mi = np.empty(max(K, 1), np.int32)
    M = L.ref_labels(C.c_int(K), _p(xyz, C.c_float), _p(mi, C.c_int32), None)
    if not want_labels:
        return mi[:M].copy(), None
    lab = np.empty((K, M), np.float32)
    L.ref_labels(C.c_int(K), _p(xyz, C.c_float), _p(mi, C.c_int32), _p(lab, C.c_float))
    return mi[:M].copy(), lab


def ref_read_poses(path, cap=1 << 20):
    L = ref_bevgen_lib()
    xyz = np.zeros((cap, 3), np.float32)
    n = L.ref_read_poses(path.encode(), _p(xyz, C.c_float), C.c_int(cap))
    return xyz[:n].copy()


def ref_list_pcd(d):
    L = ref_bevgen_lib()
    buf = C.create_string_buffer(1 << 22)
    n = L.ref_list_pcd(d.encode(), buf, C.c_int64(len(buf)))
    names = buf.value.decode().split("\n")[:-1]
    assert len(names) == n
    return names


def ref_save_labels(labels, path):
    L = ref_bevgen_lib()
    labels = np.ascontiguousarray(labels, np.float32)
    L.ref_save_labels(C.c_int(labels.shape[0]), C.c_int(labels.shape[1]), _p(labels, C.c_float), path.encode())


def ref_belonging_grid(x, y):
    L = ref_bevgen_lib()
    a, b = C.c_int(), C.c_int()
    L.ref_belonging_grid(C.c_float(x), C.c_float(y), C.byref(a), C.byref(b))
    return a.value, b.value


def ref_main(root, sensor, double_libm=False):
    """The reference's whole main() on a keyframe folder, in a child process (it prints progress and may exit()).
    Returns (returncode, stdout)."""
    import sys
    so = os.path.join(_HERE, "_ref", "libbevgen_ref_dbl.so" if double_libm else "libbevgen_ref.so")
    code = ("import ctypes as C, sys; L = C.CDLL(%r); a = [b'batch_multi_bev_gen', sys.argv[1].encode(), sys.argv[2].encode()];"
            "argv = (C.c_char_p * 4)(*a, None); rc = L.ref_bevgen_main(3, argv); sys.stdout.flush(); sys.exit(rc)" % so)
    r = subprocess.run([sys.executable, "-c", code, root, sensor], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout.decode(errors="replace"), r.stderr.decode(errors="replace")


def _ref_so(name):
    if name not in _libs:
        p = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(p):
            return None
        _libs[name] = C.CDLL(p)
    return _libs[name]


def ref_save_as_mat(x, y, z, csv_path, interval=1.0):
    """saveAsMat of the reference's own CloudManip.cpp (:79-109) -> the 201x201 f32 grid handed to imwrite; writes csv_path."""
    L = _ref_so("libcloudmanip_ref.so")
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    m = np.empty(201 * 201, np.float32)
    rc = L.ref_save_as_mat(C.c_int64(len(x)), px, py, pz, C.c_float(interval), _p(m, C.c_float), csv_path.encode())
    if rc != 201:
        raise RuntimeError("ref_save_as_mat -> %d" % rc)
    return m.reshape(201, 201)


def ref_top_flatten(x, y, z, label):
    """extractTopAndFlatten of the reference's own TopPartRegistration.cpp (:79-141) -> (out_x, out_y); None if oracle/_ref lacks it."""
    L = _ref_so("libtoppart_ref.so")
    if L is None:
        return None
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    label = np.ascontiguousarray(label, np.int16)
    n = len(x)
    ox = np.empty(max(n, 1), np.float32); oy = np.empty(max(n, 1), np.float32)
    L.ref_extract_top_and_flatten.restype = C.c_int64
    m = L.ref_extract_top_and_flatten(C.c_int64(n), px, py, pz, _p(label, C.c_int16), _p(ox, C.c_float), _p(oy, C.c_float), C.c_int64(max(n, 1)))
    if m < 0:
        raise RuntimeError("ref_extract_top_and_flatten: an output z is not 0")
    return ox[:m].copy(), oy[:m].copy()


_EXTRACTOR_SCAN = {"mulran": "sensor_data/Ouster/%010d.bin", "oxford": "velodyne_left/%010d.bin", "kitti": "velodyne/%06d.bin"}


def ref_extract_point_cloud(dataset, root, x, y, z, intensity, timestamp=1234567, double_libm=False):
    """extractPointCloud of the reference's own {Mulran,Oxford,Kitti}PointCloudSelect.cpp on a scan file written under `root` in that
    dataset's layout (MulRan / KITTI: x y z intensity per point; Oxford: all x, all y, all z, all intensities) -> dict of the
    returned cloud's fields; None if oracle/_ref lacks the build."""
    L = _ref_so("lib%sselect_ref%s.so" % (dataset, "_dbl" if double_libm else ""))
    if L is None:
        return None
    pts = [np.ascontiguousarray(a, np.float32) for a in (x, y, z, intensity)]
    path = os.path.join(root, _EXTRACTOR_SCAN[dataset] % timestamp)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    (np.stack(pts, 0) if dataset == "oxford" else np.stack(pts, 1)).tofile(path)
    cap = max(len(pts[0]) + 1, 64 * 2083)
    out = {k: np.zeros(cap, t) for k, t in (("x", np.float32), ("y", np.float32), ("z", np.float32), ("intensity", np.float32),
                                            ("row", np.uint16), ("col", np.uint16), ("label", np.int16))}
    L.ref_extract_point_cloud.restype = C.c_int64
    n = L.ref_extract_point_cloud(str(root).encode(), C.c_int64(timestamp), C.c_int64(cap), *[_p(out[k], t) for k, t in (
        ("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("intensity", C.c_float), ("row", C.c_uint16), ("col", C.c_uint16), ("label", C.c_int16))])
    assert 0 <= n <= cap, n
    return {k: v[:n] for k, v in out.items()}


def ref_cloud_manip_matrix(tx, ty, tz, theta_deg, x, y, z):
    """(rt[12], x', y', z'): the Affine3f of CloudManip.cpp:119-126 and pcl::transformPointCloud, as oracle/stub restates Eigen / PCL."""
    L = _ref_so("libcloudmanip_ref.so")
    x, px = _f(x); y, py = _f(y); z, pz = _f(z)
    rt = np.empty(12, np.float32); o = [np.empty(len(x), np.float32) for _ in range(3)]
    L.ref_cloud_manip_matrix(C.c_float(tx), C.c_float(ty), C.c_float(tz), C.c_float(theta_deg), _p(rt, C.c_float), C.c_int64(len(x)),
                             px, py, pz, *[_p(a, C.c_float) for a in o])
    return rt, o


def ref_cloud_manip_main(argv, cwd):
    """The reference's cloud_manip main() in a child process run in `cwd` (it writes its outputs into the working directory)."""
    import sys
    so = os.path.join(_HERE, "_ref", "libcloudmanip_ref.so")
    code = ("import ctypes as C, sys; L = C.CDLL(%r); a = [b'cloud_manip'] + [s.encode() for s in sys.argv[1:]];"
            "argv = (C.c_char_p * (len(a) + 1))(*a, None); rc = L.ref_cloud_manip_main(len(a), argv); sys.stdout.flush(); sys.exit(rc)" % so)
    r = subprocess.run([sys.executable, "-c", code] + list(argv), stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=cwd)
    return r.returncode, r.stdout.decode(errors="replace"), r.stderr.decode(errors="replace")


def ref_bcm_frame(x, y, z, intensity, row, col, label, out_prefix):
    """One frame through the reference's own BatchCloudManip.cpp (HDL-64E constants): (labels [S], bvm [201,201] f32)."""
    L = _ref_so("libbatchcloudmanip_ref.so")
    x, px = _f(x); y, py = _f(y); z, pz = _f(z); it, pi = _f(intensity)
    row = np.ascontiguousarray(row, np.uint16); col = np.ascontiguousarray(col, np.uint16); label = np.ascontiguousarray(label, np.int16)
    S = 64 * 2083
    lab = np.empty(S, np.int16); m = np.empty(201 * 201, np.float32)
    rc = L.ref_bcm_frame(C.c_int64(len(x)), px, py, pz, pi, _p(row, C.c_uint16), _p(col, C.c_uint16), _p(label, C.c_int16),
                         _p(lab, C.c_int16), _p(m, C.c_float), out_prefix.encode())
    if rc != S:
        raise RuntimeError("ref_bcm_frame -> %d" % rc)
    return lab, m.reshape(201, 201)


def ref_bench(sensor_name, offsets, x, y, z, intensity, row, col, label, iters=1):
    """Seconds the reference's own hot loop body (BatchMultiBevGen.cpp:735-747, PNG / CSV encoders stubbed out) spends on the
    frames, single-threaded like the reference (its globals make it non-reentrant: run one process per core)."""
    L = ref_bevgen_lib()
    L.ref_bench.restype = C.c_double
    if L.ref_set_sensor(sensor_name.encode(), None) != 0:
        raise ValueError("Unknown sensor type: %s!" % sensor_name)
    offsets = np.ascontiguousarray(offsets, np.int64)
    x, px = _f(x); y, py = _f(y); z, pz = _f(z); it, pi = _f(intensity)
    row = np.ascontiguousarray(row, np.uint16); col = np.ascontiguousarray(col, np.uint16); label = np.ascontiguousarray(label, np.int16)
    cs = C.c_uint64(0)
    return float(L.ref_bench(C.c_int(len(offsets) - 1), _p(offsets, C.c_int64), px, py, pz, pi, _p(row, C.c_uint16), _p(col, C.c_uint16),
                             _p(label, C.c_int16), C.c_int(iters), C.byref(cs)))


def ref_bench_all_cores(sensor_name, npz_path, procs, seconds=8.0):
    """frames/s of `procs` independent processes each looping ref_bench over the frames in npz_path (offsets + SoA) for about
    `seconds`.  Child processes: the reference's file-scope globals rule out threads."""
    import sys
    code = ("import sys, time, numpy as np; sys.path.insert(0, %r); import oracle_lib as O; d = np.load(sys.argv[1]); n = len(d['offsets']) - 1;"
            "a = [d[k] for k in ('x','y','z','intensity','row','col','label')]; O.ref_bench(sys.argv[2], d['offsets'], *a, iters=1);"
            "t0 = time.perf_counter(); it = 0; busy = 0.0\n"
            "while time.perf_counter() - t0 < float(sys.argv[3]): busy += O.ref_bench(sys.argv[2], d['offsets'], *a, iters=1); it += 1\n"
            "print(n * it, busy, time.perf_counter() - t0)" % _HERE)
    ps = [subprocess.Popen([sys.executable, "-c", code, npz_path, sensor_name, str(seconds)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL) for _ in range(procs)]
    frames = busy = wall = 0.0
    for p in ps:
        o = p.communicate()[0].decode().split()
        frames += float(o[0]); busy += float(o[1]); wall = max(wall, float(o[2]))
    return dict(frames_per_s=frames / wall, procs=procs, frames=int(frames), wall_s=wall, ms_per_frame_single_core=busy / frames * 1e3)
