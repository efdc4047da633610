"""Seeded synthetic keyframe generator for the three sensor shapes of src/Utility.cpp:92-124.

Frames follow the input conventions of the reference's keyframe extractors (point record
BatchMultiBevGen.h:43-53: x,y,z,intensity f32; row,col u16; t u32; label i16 = -2):
  * HDL_64E / HDL_32E: a ray-cast scene (tilted noisy ground plane, boxes, poles), ~10 % dropouts, points
    shuffled within the file, 0.5 % duplicated (row,col) slots (exercises last-writer-wins,
    BatchMultiBevGen.cpp:102-116) and 1 % `intensity = -1` markers (exercises the substitution chain :146-160).
  * OS1_64: MulRan order (row = k % 64, MulranPointCloudSelect.cpp:112-129): exactly N*H records, no-returns
    as (0,0,0), and `col` may equal Horizon_SCAN (dropped by :109).
Seed of frame i = 0x5EED0000 + i.  numpy only; no reference code involved.
"""
import numpy as np

SENSORS = {
    #            N_SCAN, Horizon_SCAN, GROUND_UPPER_SCAN, HEIGHT_RES, elev_top_deg, elev_bottom_deg
    "HDL_32E": (32, 1056, 20, 0.5, 10.67, -30.67),
    "HDL_64E": (64, 2083, 50, 0.25, 2.0, -24.8),
    "OS1_64": (64, 1024, 31, 1.0, 16.6, -16.6),
}
MAX_RANGE_M = 120.0
SENSOR_HEIGHT = 1.73


def _scene(rng):
    tilt = np.deg2rad(rng.uniform(0.0, 3.0)); tdir = rng.uniform(0, 2 * np.pi)
    n = np.array([np.sin(tilt) * np.cos(tdir), np.sin(tilt) * np.sin(tdir), np.cos(tilt)])
    boxes = []
    while len(boxes) < 40:
        r = rng.uniform(4.0, 80.0); a = rng.uniform(0, 2 * np.pi)
        cx, cy = r * np.cos(a), r * np.sin(a)
        sx, sy = rng.uniform(1.0, 15.0, 2) / 2; h = rng.uniform(1.0, 12.0)
        gap = np.hypot(max(abs(cx) - sx, 0.0), max(abs(cy) - sy, 0.0))   # footprint distance to the sensor
        if gap < 3.0:
            continue
        boxes.append((cx - sx, cx + sx, cy - sy, cy + sy, -SENSOR_HEIGHT, -SENSOR_HEIGHT + h))
    poles = []
    for _ in range(20):
        r = rng.uniform(3.0, 80.0); a = rng.uniform(0, 2 * np.pi)
        poles.append((r * np.cos(a), r * np.sin(a), rng.uniform(0.1, 0.3), -SENSOR_HEIGHT + rng.uniform(3.0, 10.0)))
    return n, boxes, poles


def _raycast(dirs, rng):
    """dirs [S,3] unit rays from the origin -> range t [S] (inf = no return)."""
    n, boxes, poles = _scene(rng)
    S = len(dirs)
    t = np.full(S, np.inf)
    dn = dirs @ n
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = -SENSOR_HEIGHT * n[2] / dn            # plane through (0,0,-h) with normal n
    tg = np.where((dn < 0) & (tg > 0), tg, np.inf)
    t = np.minimum(t, tg)
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        ix, iy, iz = 1.0 / dx, 1.0 / dy, 1.0 / dz
        for (x0, x1, y0, y1, z0, z1) in boxes:
            ta, tb = x0 * ix, x1 * ix
            tmin, tmax = np.minimum(ta, tb), np.maximum(ta, tb)
            ta, tb = y0 * iy, y1 * iy
            tmin = np.maximum(tmin, np.minimum(ta, tb)); tmax = np.minimum(tmax, np.maximum(ta, tb))
            ta, tb = z0 * iz, z1 * iz
            tmin = np.maximum(tmin, np.minimum(ta, tb)); tmax = np.minimum(tmax, np.maximum(ta, tb))
            hit = (tmax >= tmin) & (tmin > 0.5)
            t = np.where(hit & (tmin < t), tmin, t)
        a2 = dx * dx + dy * dy
        for (cx, cy, rad, ztop) in poles:
            b = dx * cx + dy * cy
            disc = b * b - a2 * (cx * cx + cy * cy - rad * rad)
            tt = (b - np.sqrt(np.maximum(disc, 0))) / a2
            zz = tt * dz
            hit = (disc > 0) & (tt > 0.5) & (zz <= ztop) & (zz >= -SENSOR_HEIGHT)
            t = np.where(hit & (tt < t), tt, t)
    t = np.where(t <= MAX_RANGE_M, t, np.inf)
    return t


def make_frame(sensor, idx, kitti_quirk=False):
    """Returns dict of SoA arrays x,y,z,intensity (f32), row,col (u16), t (u32), label (i16)."""
    N, H, G, HR, e_top, e_bot = SENSORS[sensor]
    rng = np.random.default_rng(0x5EED0000 + idx)
    elev = np.deg2rad(np.linspace(e_top, e_bot, N))                   # row 0 = top beam
    az = -np.arange(H) * (2 * np.pi / H) + rng.uniform(0, 2 * np.pi)  # clockwise sweep, random start azimuth
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    dirs = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (N, H))], -1).reshape(-1, 3)
    t = _raycast(dirs, rng)
    t = t + rng.normal(0.0, 0.03, t.shape) * np.isfinite(t)            # ±3 cm range noise
    rows, cols = np.divmod(np.arange(N * H), H)
    if sensor == "OS1_64":
        # MulRan order: k-th record has row = k % 64, column by azimuth; every record present, no-return = (0,0,0)
        k = np.arange(N * H)
        r = k % N; c = k // N
        sl = r * H + c
        tt = t[sl]
        ok = np.isfinite(tt) & (rng.random(N * H) > 0.03)
        p = dirs[sl] * np.where(ok, tt, 0.0)[:, None]
        col = c.astype(np.int64)
        bump = rng.random(N * H) < 0.002                               # round(az/360*1024) can yield 1024
        col = np.where(bump & (c == H - 1), H, col)
        out = dict(x=p[:, 0], y=p[:, 1], z=p[:, 2], intensity=rng.random(N * H) * 255.0, row=r, col=col)
    else:
        keep = np.isfinite(t) & (rng.random(N * H) > 0.10)             # ~10 % dropouts
        sl = np.nonzero(keep)[0]
        ndup = int(0.005 * len(sl))                                    # 0.5 % duplicated slots, different xyz
        dup = rng.choice(sl, ndup, replace=False)
        p = dirs[sl] * t[sl][:, None]
        pd = dirs[dup] * (t[dup] * rng.uniform(0.5, 0.99, ndup))[:, None]
        p = np.concatenate([p, pd]); r = np.concatenate([rows[sl], rows[dup]]); c = np.concatenate([cols[sl], cols[dup]])
        inten = rng.random(len(p))
        inten[rng.random(len(p)) < 0.01] = -1.0                        # 1 % "no reading" markers
        perm = rng.permutation(len(p))                                 # shuffled within the file
        out = dict(x=p[perm, 0], y=p[perm, 1], z=p[perm, 2], intensity=inten[perm], row=r[perm], col=c[perm])
        if kitti_quirk:                                                # KittiPointCloudSelect.cpp:235-240
            out["intensity"] = np.full(len(p), -1.0)
    n = len(out["x"])
    res = dict(x=out["x"].astype(np.float32), y=out["y"].astype(np.float32), z=out["z"].astype(np.float32),
               intensity=out["intensity"].astype(np.float32), row=out["row"].astype(np.uint16),
               col=out["col"].astype(np.uint16), t=np.full(n, idx, np.uint32), label=np.full(n, -2, np.int16))
    return res


def make_batch(sensor, n_frames, first=0, **kw):
    """Concatenated SoA batch: dict of arrays + 'offsets' int64[F+1]."""
    frames = [make_frame(sensor, first + i, **kw) for i in range(n_frames)]
    offs = np.zeros(n_frames + 1, np.int64)
    offs[1:] = np.cumsum([len(f["x"]) for f in frames])
    b = {k: np.concatenate([f[k] for f in frames]) for k in frames[0]}
    b["offsets"] = offs
    return b


def make_poses(n, seed=7, step=2.0, revisit=True):
    """Figure-8 trajectory of `step`-metre keyframe spacing with +-2 m z drift and one revisit.  Returns [n,3] f32."""
    rng = np.random.default_rng(seed)
    s = np.arange(n) * step
    L = max(n * step, 1.0)
    R = L / (2 * np.pi) / (2.0 if revisit else 1.0)
    th = s / R
    x = R * np.sin(th); y = R * np.sin(th) * np.cos(th)
    z = 2.0 * np.sin(2 * np.pi * s / L) + rng.normal(0, 0.05, n)
    xyz = np.stack([x, y, z], 1) + rng.normal(0, 0.2, (n, 3))
    return xyz.astype(np.float32)


def pose_csv_lines(xyz):
    """keyframe_pose.csv rows as the extractors write them (MulranPointCloudSelect.cpp:358-364):
    "{:06d}" + 15 x ",{:.6f}" = idx,x,y,z,roll,pitch,yaw,R00..R22 (identity rotation here)."""
    lines = []
    for i, (x, y, z) in enumerate(np.asarray(xyz, np.float64)):
        v = [x, y, z, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
        lines.append("%06d" % i + "".join(",%.6f" % a for a in v))
    return lines


def make_kitti_scan(seed, n_rings=64, start_negative=False, short_rings=(17, 40), jitter=True):
    """Raw HDL-64E scan in KITTI file order (ring after ring, every ring sweeping the azimuth 0+ .. 180, -180 .. 0-),
    for the ring detection of KittiPointCloudSelect.cpp:188-243: rings in `short_rings` hold fewer than 1250 points
    (their crossing must be ignored, :217), `jitter` adds sign flips right after a ring start (spurious crossings),
    `start_negative` begins below 0 degrees (ring_idx starts at -1, :199-204)."""
    rng = np.random.default_rng(0xC17 + seed)
    az_all = []
    if start_negative:
        az_all.append(np.sort(rng.uniform(-20.0, -0.5, 37)))
    for r in range(n_rings):
        m = int(rng.integers(700, 1100)) if r in short_rings else int(rng.integers(1500, 2084))
        a = np.sort(rng.uniform(0.05, 359.95, m))
        a = np.where(a > 180.0, a - 360.0, a)
        if jitter and r % 5 == 2:
            a[1:6] = [-0.3, 0.2, -0.1, 0.4, 0.6]          # flips right after the ring start: too few points, ignored
        az_all.append(a)
    az = np.concatenate(az_all)
    rad = rng.uniform(3.0, 70.0, len(az))
    x = (rad * np.cos(np.deg2rad(az))).astype(np.float32); y = (rad * np.sin(np.deg2rad(az))).astype(np.float32)
    z = rng.uniform(-2.0, 1.0, len(az)).astype(np.float32)
    return x, y, z
