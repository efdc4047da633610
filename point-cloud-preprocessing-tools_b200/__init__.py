"""point-cloud-preprocessing-tools_b200 — B200 (sm_100a) implementation of the `batch_multi_bev_gen` hot path of
soytony/Point-Cloud-Preprocessing-Tools (BatchMultiBevGen.cpp).

The product is the C-ABI library lib/libbevgen_cuda.so (include/bevgen.h) and the C++ CLIs in bin/.  This module is
only a thin ctypes mirror of that C-ABI for tests and bench.py: it holds no compute of its own and it fails loudly
when the CUDA library is missing or unusable — there is no CPU fallback.

The directory name is not a Python identifier; load it with `_load_pkg.py` at the repo root (alias `pcpt_b200`).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbevgen_cuda.so")
CLI_PATH = os.path.join(_HERE, "bin", "batch_multi_bev_gen")
CLOUD_MANIP_PATH = os.path.join(_HERE, "bin", "cloud_manip")
BATCH_CLOUD_MANIP_PATH = os.path.join(_HERE, "bin", "batch_cloud_manip")
GRID, LAYERS, CELLS = 224, 24, 224 * 224
N_STAGES = 8


class Params(C.Structure):
    _fields_ = [("n_scan", C.c_int32), ("horizon_scan", C.c_int32), ("ground_upper_scan", C.c_int32),
                ("height_res", C.c_float), ("grid_size", C.c_int32), ("max_range", C.c_int32), ("n_layers", C.c_int32),
                ("lidar_to_ground", C.c_float), ("rt", C.c_float * 12), ("has_transform", C.c_int32)]

    @property
    def S(self):
        return self.n_scan * self.horizon_scan


class Points(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("intensity", C.c_void_p),
                ("row", C.c_void_p), ("col", C.c_void_p), ("label", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("label", C.c_void_p), ("winner_bits", C.c_void_p), ("single_bev", C.c_void_p), ("multi_bev", C.c_void_p),
                ("bvm", C.c_void_p)]     # bvm: optional [F][201][201] f32 bird-view map (batch_cloud_manip), None = skip


class RecordLayout(C.Structure):
    """bevgen_record_layout: byte offsets of the PointXYZIRCT fields inside an interleaved record (-1 = absent)."""
    _fields_ = [("stride", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_row", C.c_int32), ("off_col", C.c_int32), ("off_label", C.c_int32)]


class PointsCompact(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("meta", C.c_void_p)]


class OutputsCompact(C.Structure):
    _fields_ = [("ground_bits", C.c_void_p), ("winner_bits", C.c_void_p), ("single_bev", C.c_void_p), ("multi_planes", C.c_void_p)]


META_INVALID, META_NEG1, META_LABELED = 0x00FFFFFF, 1 << 24, 1 << 25


def pack_meta(params, row, col, intensity, label):
    """bevgen_pack_meta (include/bevgen.h) over arrays: slot | intensity == -1 | label != 0 - a change of representation."""
    row = np.asarray(row).astype(np.int64); col = np.asarray(col).astype(np.int64)
    ok = (row < params.n_scan) & (col < params.horizon_scan)
    slot = np.where(ok, row * params.horizon_scan + col, META_INVALID).astype(np.uint32)
    return slot | np.where(np.asarray(intensity, np.float32) == np.float32(-1.0), np.uint32(META_NEG1), np.uint32(0)) \
        | np.where(np.asarray(label) != 0, np.uint32(META_LABELED), np.uint32(0))


def expand_multi(planes):
    """bevgen_expand_multi over frames: [F][3][224][224] bit planes -> [F][24][224][224] 0/255 layers."""
    planes = np.asarray(planes, np.uint8)
    l = np.arange(LAYERS)
    return (((planes[:, l >> 3] >> (l & 7)[None, :, None, None].astype(np.uint8)) & 1) * 255).astype(np.uint8)


def labels_from_ground_bits(ground_bits, owner, label_in, offsets, frames=None):
    """[F][S] int16 labels of the ordered cloud from the compact outputs: 0 where the slot is ground (or empty), else the
    label of the input point that owns the slot (BatchMultiBevGen.cpp:244-248 leaves it untouched).  frames: the frame
    indices the rows of `owner` stand for (default: all frames of the batch)."""
    F = len(offsets) - 1
    frames = list(range(F)) if frames is None else list(frames)
    S = owner.shape[1]
    gb = np.asarray(ground_bits, np.uint32).reshape(F, -1)
    lab = np.zeros((len(frames), S), np.int16)
    for j, f in enumerate(frames):
        g = np.unpackbits(gb[f].view(np.uint8), bitorder="little")[:S].astype(bool)
        sel = (owner[j] > 0) & ~g
        lab[j, sel] = np.asarray(label_in)[int(offsets[f]) + owner[j][sel].astype(np.int64) - 1]
    return lab


def pcd_record_layout():
    """The 26-byte record of a PointXYZIRCT binary PCD (savePCDFileBinary, BatchMultiBevGen.cpp:756)."""
    lay = RecordLayout()
    _ck(lib().bevgen_pcd_record_layout(C.byref(lay)))
    return lay


def winner_words(n_total, n_frames):
    """bevgen_winner_words (include/bevgen.h)."""
    return (int(n_total) >> 5) + int(n_frames) + 1


def winner_mask(winner, offsets, f):
    """bool[n_f]: which input points of frame f survive getOrderedCloud (unpacks bevgen_outputs.winner_bits)."""
    o, e = int(offsets[f]), int(offsets[f + 1])
    n = e - o
    w0 = (o >> 5) + f
    words = np.asarray(winner[w0:w0 + (n + 31) // 32], np.uint32)
    return np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)


def owner_from_winner(winner, offsets, row, col, H, S, frames=None):
    """[F][S] uint32, 1 + index of the input point that occupies the slot (0 = empty): the ordered cloud of
    getOrderedCloud (BatchMultiBevGen.cpp:94-117) expressed as a gather table, rebuilt on the host from the winner bits
    and the caller's own row/col — what a host does to write non_ground_point_cloud/*.pcd (:756)."""
    frames = list(range(len(offsets) - 1)) if frames is None else list(frames)
    own = np.zeros((len(frames), S), np.uint32)
    for j, f in enumerate(frames):
        o, e = int(offsets[f]), int(offsets[f + 1])
        idx = np.nonzero(winner_mask(winner, offsets, f))[0]
        slot = row[o:e][idx].astype(np.int64) * H + col[o:e][idx].astype(np.int64)
        own[j, slot] = (idx + 1).astype(np.uint32)
    return own


EXPORTS = ["bevgen_sensor_params", "bevgen_create", "bevgen_destroy", "bevgen_last_error", "bevgen_host_alloc",
           "bevgen_host_free", "bevgen_process_host", "bevgen_process_device", "bevgen_sync", "bevgen_submit",
           "bevgen_collect", "bevgen_select_major", "bevgen_labels", "bevgen_cloud_manip", "bevgen_set_profiling",
           "bevgen_stage_ms", "bevgen_kernel_launches", "bevgen_compute_stream", "bevgen_stage_name",
           "bevgen_debug_atan2f", "bevgen_pcd_record_layout", "bevgen_process_packed_host", "bevgen_project", "bevgen_top_flatten",
           "bevgen_process_host_compact", "bevgen_cloud_manip_device", "bevgen_set_libm", "bevgen_set_diag", "bevgen_get_diag",
           "bevgen_host_alloc_wc"]


def build(verbose=False):
    """make -C <package>: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=out)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libbevgen_cuda.so is not built (%s); run __graft_entry__.build() — there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.bevgen_last_error.restype = C.c_char_p
        L.bevgen_stage_name.restype = C.c_char_p
        L.bevgen_host_alloc.restype = C.c_void_p
        L.bevgen_host_alloc.argtypes = [C.c_size_t]
        L.bevgen_host_alloc_wc.restype = C.c_void_p
        L.bevgen_host_alloc_wc.argtypes = [C.c_size_t]
        L.bevgen_host_free.argtypes = [C.c_void_p]
        L.bevgen_kernel_launches.restype = C.c_int64
        L.bevgen_kernel_launches.argtypes = [C.c_void_p]
        L.bevgen_compute_stream.restype = C.c_void_p
        L.bevgen_compute_stream.argtypes = [C.c_void_p]
        L.bevgen_destroy.argtypes = [C.c_void_p]
        L.bevgen_destroy.restype = None
        _lib = L
    return _lib


class BevgenError(RuntimeError):
    pass


def _ck(rc):
    if rc < 0:
        raise BevgenError(lib().bevgen_last_error().decode())
    return rc


def sensor_params(name):
    p = Params()
    rc = lib().bevgen_sensor_params(name.encode(), C.byref(p))
    if rc < 0:
        raise BevgenError(lib().bevgen_last_error().decode())
    return p


def _ptr(a):
    """host numpy array or integer device pointer -> void*"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype, write_combined=False):
    """numpy array backed by cudaHostAlloc memory (release with pinned_free).  write_combined: input staging only - the host
    must never read such an array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = (lib().bevgen_host_alloc_wc if write_combined else lib().bevgen_host_alloc)(max(n, 1))
    if not p:
        raise BevgenError(lib().bevgen_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.__array_interface__["data"][0], None)
    if p:
        lib().bevgen_host_free(C.c_void_p(p))


class BevGen:
    """One context = one GPU (bevgen_ctx).  Mirrors the hot loop body of BatchMultiBevGen.cpp:727-757."""

    def __init__(self, sensor="HDL_64E", device=0, max_points_per_frame=None, max_frames_per_batch=64, rt=None):
        self.params = sensor_params(sensor) if isinstance(sensor, str) else sensor
        if rt is not None:
            rt = np.ascontiguousarray(rt, np.float32).reshape(12)
            for i in range(12):
                self.params.rt[i] = float(rt[i])
            self.params.has_transform = 1
        self.S = self.params.S
        self.max_pts = int(max_points_per_frame or self.S + 4096)
        self.max_frames = int(max_frames_per_batch)
        self._ctx = C.c_void_p()
        _ck(lib().bevgen_create(C.byref(self._ctx), C.c_int(device), C.byref(self.params), C.c_int(self.max_pts),
                                C.c_int(self.max_frames)))

    def close(self):
        if self._ctx:
            lib().bevgen_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hot loop -------------------------------------------------------------------------------------------
    def alloc_outputs(self, F, pinned=False, n_total=None, bvm=False):
        mk = pinned_empty if pinned else np.empty
        n_total = F * self.max_pts if n_total is None else n_total
        out = dict(label=mk((F, self.S), np.int16), winner=mk((winner_words(n_total, F),), np.uint32),
                   single=mk((F, GRID, GRID), np.uint8), multi=mk((F, LAYERS, GRID, GRID), np.uint8))
        if bvm:
            out["bvm"] = mk((F, 201, 201), np.float32)
        out["winner"][...] = 0      # the library writes every word of a frame's range and zeroes the gaps; the tail word is spare
        return out

    def process_host(self, batch, out=None):
        """batch: dict x,y,z,intensity,row,col,label (+offsets int64[F+1]) of HOST numpy arrays."""
        offs = np.ascontiguousarray(batch["offsets"], np.int64)
        F = len(offs) - 1
        user_out = out is not None
        out = out or self.alloc_outputs(F, n_total=int(offs[-1]))
        arrs = [_as(batch[k], t) for k, t in _FIELDS]      # keep converted temporaries alive across the call
        pts = Points(*[_ptr(a) for a in arrs])
        o = Outputs(_ptr(out["label"]), _ptr(out["winner"]), _ptr(out["single"]), _ptr(out["multi"]), _ptr(out.get("bvm")))
        _ck(lib().bevgen_process_host(self._ctx, C.c_int(F), _ptr(offs), C.byref(pts), C.byref(o)))
        if not user_out:   # convenience for tests: the ordered cloud as a gather table (host-side unpack of the winner bits)
            out["owner"] = owner_from_winner(out["winner"], offs, arrs[4], arrs[5], self.params.horizon_scan, self.S)
        return out

    def alloc_outputs_compact(self, F, pinned=False, n_total=None):
        mk = pinned_empty if pinned else np.empty
        n_total = F * self.max_pts if n_total is None else n_total
        out = dict(ground=mk((F, (self.S + 31) // 32), np.uint32), winner=mk((winner_words(n_total, F),), np.uint32),
                   single=mk((F, GRID, GRID), np.uint8), planes=mk((F, 3, GRID, GRID), np.uint8))
        out["winner"][...] = 0
        return out

    def process_host_compact(self, cbatch, out=None):
        """cbatch: dict x, y, z (f32), meta (u32, pack_meta) + offsets of HOST arrays -> compact outputs (ground, winner, single, planes)."""
        offs = np.ascontiguousarray(cbatch["offsets"], np.int64)
        F = len(offs) - 1
        out = out or self.alloc_outputs_compact(F, n_total=int(offs[-1]))
        arrs = [_as(cbatch[k], t) for k, t in (("x", np.float32), ("y", np.float32), ("z", np.float32), ("meta", np.uint32))]
        pts = PointsCompact(*[_ptr(a) for a in arrs])
        o = OutputsCompact(_ptr(out["ground"]), _ptr(out["winner"]), _ptr(out["single"]), _ptr(out["planes"]))
        _ck(lib().bevgen_process_host_compact(self._ctx, C.c_int(F), _ptr(offs), C.byref(pts), C.byref(o)))
        return out

    def compact_to_reference_layout(self, cout, batch, frames=None):
        """Expands compact outputs into the dict process_host returns (owner, label, single, multi), using the caller's
        own row / col / label arrays - what the CLI's encode pool does before writing files.  frames: only these frame
        indices (rows of the result follow that list)."""
        offs = np.asarray(batch["offsets"], np.int64)
        owner = owner_from_winner(cout["winner"], offs, _as(batch["row"], np.uint16), _as(batch["col"], np.uint16), self.params.horizon_scan, self.S, frames)
        sel = slice(None) if frames is None else list(frames)
        return dict(owner=owner, label=labels_from_ground_bits(cout["ground"], owner, batch["label"], offs, frames),
                    single=np.asarray(cout["single"])[sel], multi=expand_multi(np.asarray(cout["planes"])[sel]))

    def process_packed_host(self, records, offsets, layout=None, out=None):
        """records: HOST uint8 array with the concatenated interleaved records of all frames (a binary PCD payload);
        layout: RecordLayout (default = the 26-byte PCD record).  The de-interleave runs on the GPU."""
        offs = np.ascontiguousarray(offsets, np.int64)
        F = len(offs) - 1
        layout = layout or pcd_record_layout()
        rec = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
        if rec.size < int(offs[-1]) * layout.stride:
            raise ValueError("records shorter than offsets[-1] * stride")
        out = out or self.alloc_outputs(F, n_total=int(offs[-1]))
        o = Outputs(_ptr(out["label"]), _ptr(out["winner"]), _ptr(out["single"]), _ptr(out["multi"]), _ptr(out.get("bvm")))
        _ck(lib().bevgen_process_packed_host(self._ctx, C.c_int(F), _ptr(offs), _ptr(rec), C.byref(layout), C.byref(o)))
        return out

    def process_device(self, F, offsets, dev_in, dev_out):
        """dev_in / dev_out: dicts of integer DEVICE pointers (same keys as the host form); async on the compute stream."""
        offs = np.ascontiguousarray(offsets, np.int64)
        pts = Points(*[C.c_void_p(int(dev_in[k])) for k, _ in _FIELDS])
        o = Outputs(C.c_void_p(int(dev_out["label"])), C.c_void_p(int(dev_out["winner"])), C.c_void_p(int(dev_out["single"])),
                    C.c_void_p(int(dev_out["multi"])), C.c_void_p(int(dev_out["bvm"])) if dev_out.get("bvm") else None)
        _ck(lib().bevgen_process_device(self._ctx, C.c_int(F), _ptr(offs), C.byref(pts), C.byref(o)))

    def sync(self):
        _ck(lib().bevgen_sync(self._ctx))

    def submit(self, frame_id, f):
        n = len(f["x"])
        a = [_as(f[k], t) for k, t in _FIELDS]
        _ck(lib().bevgen_submit(self._ctx, C.c_int(frame_id), C.c_int(n), *[_ptr(v) for v in a]))

    def collect(self, frame_id, frame=None):
        """frame: the submitted frame (for row/col) - then the result also carries the 'owner' gather table."""
        out = self.alloc_outputs(1)
        _ck(lib().bevgen_collect(self._ctx, C.c_int(frame_id), _ptr(out["label"]), _ptr(out["winner"]), _ptr(out["single"]),
                                 _ptr(out["multi"])))
        res = {k: (v if k == "winner" else v[0]) for k, v in out.items()}
        if frame is not None:
            n = len(frame["row"])
            res["owner"] = owner_from_winner(out["winner"], np.array([0, n]), _as(frame["row"], np.uint16), _as(frame["col"], np.uint16),
                                             self.params.horizon_scan, self.S)[0]
        return res

    # ---- labels ---------------------------------------------------------------------------------------------
    def select_major(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        K = len(xyz)
        mi = np.empty(max(K, 1), np.int32); ov = np.empty(max(K, 1), np.int32); M = C.c_int32(0)
        _ck(lib().bevgen_select_major(self._ctx, C.c_int(K), _ptr(xyz), _ptr(mi), C.byref(M), _ptr(ov)))
        return mi[:M.value].copy(), ov[:K].copy()

    def labels(self, xyz, major_idx, row_begin=0, row_end=None, dense=True):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        major_idx = np.ascontiguousarray(major_idx, np.int32)
        K, M = len(xyz), len(major_idx)
        row_end = K if row_end is None else row_end
        rows = row_end - row_begin
        lab = np.empty((rows, M), np.float32) if dense else None
        nn = np.empty((rows, 2), np.int32); w = np.empty((rows, 2), np.float32)
        _ck(lib().bevgen_labels(self._ctx, C.c_int(K), _ptr(xyz), C.c_int(M), _ptr(major_idx), C.c_int(row_begin),
                                C.c_int(row_end), _ptr(lab), _ptr(nn), _ptr(w)))
        return lab, nn, w

    # ---- cloud_manip ------------------------------------------------------------------------------------------
    def cloud_manip(self, rt, x, y, z):
        rt = np.ascontiguousarray(rt, np.float32).reshape(12)
        x = _as(x, np.float32); y = _as(y, np.float32); z = _as(z, np.float32)
        n = len(x)
        t = [np.empty(n, np.float32) for _ in range(3)]
        bi = np.empty((201, 201), np.float32); bo = np.empty((201, 201), np.float32)
        _ck(lib().bevgen_cloud_manip(self._ctx, C.c_int64(n), _ptr(rt), _ptr(x), _ptr(y), _ptr(z), _ptr(t[0]), _ptr(t[1]),
                                     _ptr(t[2]), _ptr(bi), _ptr(bo)))
        return t, bi, bo

    def cloud_manip_device(self, n, rt, dev):
        """dev: dict of integer DEVICE pointers x, y, z, tx, ty, tz, bev_in, bev_out (outputs may be 0); async on the compute stream."""
        rt = np.ascontiguousarray(rt, np.float32).reshape(12)
        g = lambda k: C.c_void_p(int(dev[k])) if dev.get(k) else None
        _ck(lib().bevgen_cloud_manip_device(self._ctx, C.c_int64(n), _ptr(rt), g("x"), g("y"), g("z"), g("tx"), g("ty"), g("tz"), g("bev_in"), g("bev_out")))

    # ---- libm overload set / diagnostics of the ground criterion ------------------------------------------------
    def set_libm(self, use_double):
        _ck(lib().bevgen_set_libm(self._ctx, C.c_int(1 if use_double else 0)))

    def set_diag(self, on):
        _ck(lib().bevgen_set_diag(self._ctx, C.c_int(1 if on else 0)))

    def get_diag(self):
        out = (C.c_uint64 * 4)()
        _ck(lib().bevgen_get_diag(self._ctx, out))
        return dict(borderline_pairs=int(out[0]), float_double_disagree=int(out[1]))

    # ---- projection step of the keyframe extractors -----------------------------------------------------------
    def project(self, kind, x, y, z=None):
        """kind: 0 = MulRan OS1-64 (row = k % 64), 1 = Oxford HDL-32E (returns the negated x, z too), 2 = KITTI HDL-64E ring
        detection (row = col = 0xFFFF for points the extractor does not place).  -> dict."""
        x = np.array(x, np.float32); y = _as(y, np.float32)
        z = None if z is None else np.array(z, np.float32)
        n = len(x)
        row = np.empty(n, np.uint16); col = np.empty(n, np.uint16)
        _ck(lib().bevgen_project(self._ctx, C.c_int(kind), C.c_int64(n), _ptr(x), _ptr(y), _ptr(z), _ptr(row), _ptr(col)))
        return dict(x=x, y=y, z=z, row=row, col=col)

    # ---- extractTopAndFlatten -----------------------------------------------------------------------------------
    def top_flatten(self, x, y, z, label):
        """TopPartRegistration.cpp:79-141 -> (out_x, out_y, source_index)."""
        x = _as(x, np.float32); y = _as(y, np.float32); z = _as(z, np.float32); label = _as(label, np.int16)
        n = len(x)
        ox = np.empty(max(n, 1), np.float32); oy = np.empty(max(n, 1), np.float32); oi = np.empty(max(n, 1), np.uint32)
        m = C.c_int64(0)
        _ck(lib().bevgen_top_flatten(self._ctx, C.c_int64(n), _ptr(x), _ptr(y), _ptr(z), _ptr(label), _ptr(ox), _ptr(oy), _ptr(oi), C.byref(m)))
        return ox[:m.value].copy(), oy[:m.value].copy(), oi[:m.value].copy()

    # ---- introspection ----------------------------------------------------------------------------------------
    def set_profiling(self, on):
        _ck(lib().bevgen_set_profiling(self._ctx, C.c_int(1 if on else 0)))

    def stage_ms(self):
        ms = (C.c_float * N_STAGES)(); ln = (C.c_int64 * N_STAGES)()
        _ck(lib().bevgen_stage_ms(self._ctx, ms, ln))
        return {lib().bevgen_stage_name(i).decode(): (ms[i], ln[i]) for i in range(N_STAGES)}

    def kernel_launches(self):
        return int(lib().bevgen_kernel_launches(self._ctx))

    def compute_stream(self):
        return int(lib().bevgen_compute_stream(self._ctx) or 0)

    def debug_atan2f(self, y, x):
        y = _as(y, np.float32); x = _as(x, np.float32)
        out = np.empty(len(y), np.float32)
        _ck(lib().bevgen_debug_atan2f(self._ctx, C.c_int64(len(y)), _ptr(y), _ptr(x), _ptr(out)))
        return out


_FIELDS = [("x", np.float32), ("y", np.float32), ("z", np.float32), ("intensity", np.float32), ("row", np.uint16),
           ("col", np.uint16), ("label", np.int16)]


def _as(a, t):
    a = np.asarray(a)
    if a.dtype != t or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, t)
    return a
