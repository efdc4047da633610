"""Frame constructors shared by the CPU pin tests (oracle vs the reference's own compiled source, oracle/_ref) and the
GPU parity tests (CUDA vs oracle): the same inputs prove both links of the chain reference -> oracle -> CUDA."""
import numpy as np

FIELD_TYPES = (("x", np.float32), ("y", np.float32), ("z", np.float32), ("intensity", np.float32),
               ("row", np.uint16), ("col", np.uint16), ("label", np.int16))


def rand_frame(rng, N, H, n, spread=60.0, zlo=-3.0, zhi=6.0, p_neg1=0.05, col_over=True):
    """Uniform random points: slot collisions, out-of-range row/col, mixed labels incl. 0, -1 intensities."""
    f = dict(x=rng.uniform(-spread, spread, n), y=rng.uniform(-spread, spread, n), z=rng.uniform(zlo, zhi, n),
             intensity=np.where(rng.random(n) < p_neg1, -1.0, rng.random(n)),
             row=rng.integers(0, N + (2 if col_over else 0), n), col=rng.integers(0, H + (3 if col_over else 0), n),
             label=rng.integers(-3, 4, n))
    return {k: np.asarray(f[k]).astype(t) for k, t in FIELD_TYPES}


def random_unstructured_frames(sp, seed=1234):
    """The frame list of test_random_unstructured_frames (heavy collisions, tiny / empty frames, wide spread)."""
    rng = np.random.default_rng(seed)
    frames = [rand_frame(rng, sp.n_scan, sp.horizon_scan, n) for n in (sp.S * 2, sp.S // 3, 1, 0, 5000)]
    frames.append(rand_frame(rng, sp.n_scan, sp.horizon_scan, sp.S, spread=130.0, zlo=-10, zhi=30, p_neg1=0.5))
    return frames


def borderline_frames(sp, seed=7):
    """Two organised frames whose vertical neighbours sit within +-40 ulp of the 10-degree threshold
    (BatchMultiBevGen.cpp:173-179); the second also carries inf / nan / huge / denormal coordinates."""
    N, H = sp.n_scan, sp.horizon_scan
    rng = np.random.default_rng(seed)
    rows, cols = np.divmod(np.arange(N * H), H)
    rng_h = rng.uniform(0.5, 40.0, N * H).astype(np.float32)
    x = np.zeros(N * H, np.float32); y = np.zeros(N * H, np.float32); z = np.zeros(N * H, np.float32)
    tan10 = np.tan(np.float64(0.17453292))
    for r in range(N - 1, -1, -1):        # build columns bottom-up so that dz/h of (r-1, r) is ~tan(10 deg) * (1 + k ulp)
        sel = rows == r
        if r == N - 1:
            x[sel] = rng.uniform(-30, 30, H); y[sel] = rng.uniform(-30, 30, H); z[sel] = -1.7
        else:
            below = rows == r + 1
            h = rng_h[sel]
            ang = rng.uniform(0, 2 * np.pi, H)
            x[sel] = x[below] + (h * np.cos(ang)).astype(np.float32)
            y[sel] = y[below] + (h * np.sin(ang)).astype(np.float32)
            dx = x[sel] - x[below]; dy = y[sel] - y[below]
            hh = np.sqrt((dx * dx + dy * dy).astype(np.float32)).astype(np.float64)
            k = rng.integers(-40, 41, H)
            sgn = np.where(rng.random(H) < 0.5, -1.0, 1.0)
            z[sel] = z[below] + (sgn * hh * tan10 * (1.0 + k * 6e-8)).astype(np.float32)
    f = dict(x=x, y=y, z=z, intensity=np.full(N * H, 0.5, np.float32), row=rows.astype(np.uint16),
             col=cols.astype(np.uint16), label=np.full(N * H, -2, np.int16))
    f2 = {k: v.copy() for k, v in f.items()}
    idx = rng.choice(N * H, 600, replace=False)
    f2["z"][idx[:100]] = np.inf; f2["x"][idx[100:200]] = np.nan; f2["x"][idx[200:300]] = 3e38
    f2["z"][idx[300:400]] = -np.inf; f2["y"][idx[400:500]] = -3e38; f2["z"][idx[500:600]] = 1e-42
    return [f, f2]


def boundary_frame(sp):
    """Cell / layer / height rounding at the boundaries listed in SURVEY §8a.1-B (intensity -1: nothing is ground)."""
    vs = np.array([-113.0, -112.99999, -112.5, -112.0, -111.99999, -111.5, -111.0, 0.0, -0.0, 0.49999997, 0.5, 110.99999,
                   111.0, 111.00001, 110.5, 112.0, 1e9, -1e9, np.nan, np.inf], np.float32)
    zs = np.array([-2.0, -1.26, -1.25, -1.24999, -1.0, -0.75, -0.7500001, 0.0, 10.74, 10.75, 10.76, 11.0, 61.7, 61.75, 70.0,
                   -2.1, 5.3e8, 6e8, np.nan, -np.inf, np.inf, 1e-40], np.float32)
    X, Y, Z = np.meshgrid(vs, vs, zs, indexing="ij")
    n = X.size
    assert n <= sp.S
    slots = np.arange(n)
    return dict(x=X.ravel(), y=Y.ravel(), z=Z.ravel(), intensity=np.full(n, -1.0, np.float32),
                row=(slots // sp.horizon_scan).astype(np.uint16), col=(slots % sp.horizon_scan).astype(np.uint16),
                label=np.where(slots % 7 == 0, 0, 5).astype(np.int16))


def hot_cell_frame(sp, seed=3, n=None, jitter=0.3):
    """Contention stress: every point of a full frame falls into ONE 1 m BEV cell (and one 2 m ground sector):
    the per-point read-modify-write of BatchMultiBevGen.cpp:289-291, 353-355 at its worst."""
    rng = np.random.default_rng(seed)
    n = sp.S if n is None else n
    slots = rng.permutation(sp.S)[:n] if n <= sp.S else rng.integers(0, sp.S, n)
    return dict(x=(10.2 + rng.uniform(0, jitter, n)).astype(np.float32), y=(-7.6 + rng.uniform(0, jitter, n)).astype(np.float32),
                z=rng.uniform(-1.9, 3.9, n).astype(np.float32), intensity=rng.random(n).astype(np.float32),
                row=(slots // sp.horizon_scan).astype(np.uint16), col=(slots % sp.horizon_scan).astype(np.uint16),
                label=np.full(n, -2, np.int16))


# ---- golden vectors generated from the reference's own source (tests/golden/make_bev_golden.py) ------------------
GOLDEN_LABEL_SETS = ((100, 5, 9.0), (1500, 8, 2.0), (10000, 11, 2.0), (30, 4, 0.1))   # (K, seed, step): M = 11, 67, 457, 1


def golden_frame_list(synth, O):
    """(case id, sensor, frame dict): the enumeration shared by the fixture generator and the tests."""
    out = []
    for sensor in ("HDL_32E", "OS1_64", "HDL_64E"):
        sp = O.sensor(sensor)
        for idx in (0, 1, 300):
            out.append(("synth/%s/%d" % (sensor, idx), sensor, synth.make_frame(sensor, idx)))
        for i, f in enumerate(random_unstructured_frames(sp)):
            out.append(("random/%s/%d" % (sensor, i), sensor, f))
        out.append(("hot/%s" % sensor, sensor, hot_cell_frame(sp)))
    sp = O.sensor("HDL_32E")
    for i, f in enumerate(borderline_frames(sp)):
        out.append(("borderline/%d" % i, "HDL_32E", f))
    out.append(("boundary", "HDL_32E", boundary_frame(sp)))
    out.append(("kitti_quirk", "HDL_64E", synth.make_frame("HDL_64E", 100, kitti_quirk=True)))
    return out


def digest(a):
    import hashlib
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:                       # NaN payloads are not part of the contract
        a = np.where(np.isnan(a), np.float32(np.nan), a).view(np.uint32)
    return hashlib.sha256(a.tobytes()).hexdigest()


def top_flatten_cases(O, synth):
    """Inputs of extractTopAndFlatten: a real ground-removed keyframe (what non_ground_point_cloud/*.pcd holds), wide random
    clouds with distinct heights, the 20-point threshold, the cell borders, an empty cloud.  The reference sorts a cell's
    points with std::sort, whose order of EQUAL heights is unspecified, so every case but the last keeps heights distinct."""
    sp = O.sensor("HDL_64E")
    f = synth.make_frame("HDL_64E", 77)
    oc = O.order(sp, *[f[k] for k, _ in FIELD_TYPES])
    lab = O.mark_ground(sp, oc)[0]
    zz = oc["z"].copy(); zz += np.arange(len(zz), dtype=np.float32) * np.float32(1e-7)      # break the (many) equal heights of empty slots
    out = [("keyframe", oc["x"], oc["y"], zz.astype(np.float32), lab)]
    rng = np.random.default_rng(21)
    for n, spread in ((300_007, 130.0), (5_003, 60.0), (977, 12.0)):
        z = rng.permutation(n).astype(np.float32) * np.float32(0.001) - np.float32(3.0)      # distinct
        out.append(("random %d" % n, rng.uniform(-spread, spread, n).astype(np.float32), rng.uniform(-spread, spread, n).astype(np.float32),
                    z, rng.integers(-2, 3, n).astype(np.int16)))
    # exactly 19 / 20 / 22 / 23 points in a cell (:124 threshold, :123 round(0.2 * n): 4, 4, 5), cell borders at +-10 m (:104-105 round)
    xs, ys, zs = [], [], []
    for cx, cnt in ((-95.0, 19), (-75.0, 20), (-55.0, 22), (-35.0, 23), (9.999, 30), (10.0, 30), (10.001, 30), (-100.0, 25), (99.999, 25), (100.0, 25)):
        xs += [cx] * cnt; ys += [5.0] * cnt; zs += list(np.arange(cnt) * 0.5 + len(zs))
    out.append(("thresholds", np.array(xs, np.float32), np.array(ys, np.float32), np.array(zs, np.float32), np.ones(len(xs), np.int16)))
    out.append(("empty", np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.int16)))
    return out


KITTI_SCANS = [(0, {}), (1, dict(start_negative=True)), (2, dict(n_rings=70, short_rings=(0, 1, 33))), (3, dict(n_rings=3, jitter=False))]
KITTI_H = 2083                                      # KittiPointCloudSelect.cpp:148-149
PROJECTION_SPECIALS = [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e30, np.inf, -np.inf, np.nan, 1e-7, -1e-7]


def projection_cloud(n=300_001, seed=11):
    """The cloud of test_projection_step_bit_exact: normal coordinates with every pair of special values (zeros of both signs,
    axis points, huge / tiny / non-finite) in the first 144 points."""
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 30, n).astype(np.float32); y = rng.normal(0, 30, n).astype(np.float32); z = rng.normal(-1, 3, n).astype(np.float32)
    sp = PROJECTION_SPECIALS
    k = 0
    for a in sp:
        for b in sp:
            x[k], y[k], z[k] = a, b, sp[(k * 7) % len(sp)]; k += 1
    return x, y, z


def kitti_structured(x, y, z, row, col):
    """The structured cloud KittiPointCloudSelect.cpp:206-243 returns, rebuilt from per-point (row, col) (0xFFFF = not placed):
    slot row * 2083 + col holds its LAST writer with intensity -1 and label -2 (:233-238), every other slot is all-zero (:207)."""
    S = 64 * KITTI_H
    out = {k: np.zeros(S, t) for k, t in (("x", np.float32), ("y", np.float32), ("z", np.float32), ("intensity", np.float32),
                                          ("row", np.uint16), ("col", np.uint16), ("label", np.int16))}
    placed = np.nonzero(np.asarray(row) != 0xFFFF)[0]
    slot = np.asarray(row)[placed].astype(np.int64) * KITTI_H + np.asarray(col)[placed]
    last = np.full(S, -1, np.int64)
    np.maximum.at(last, slot, placed)                # the serial loop's last writer = the largest input index
    s = np.nonzero(last >= 0)[0]; i = last[s]
    out["x"][s] = x[i]; out["y"][s] = y[i]; out["z"][s] = z[i]
    out["intensity"][s] = -1; out["row"][s] = np.asarray(row)[i]; out["col"][s] = np.asarray(col)[i]; out["label"][s] = -2
    return out


def structured_digest(c):
    return digest(np.concatenate([np.asarray(c[k]).view(np.uint8) if c[k].dtype != np.float32 else
                                  np.where(np.isnan(c[k]), np.float32(np.nan), c[k]).view(np.uint8)
                                  for k in ("x", "y", "z", "intensity", "row", "col", "label")]))
