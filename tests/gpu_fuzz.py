#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it loads the oracle).  Seeded fuzz of the CUDA path against the oracle beyond the fixed frames of the
parity tests (a checker run, not a benchmark):
organised "scene" frames with random ground planes / walls / dropouts / -1 markers, random unstructured frames of random size,
hot-cell frames, all three sensors, through bevgen_process_host AND bevgen_process_host_compact (after host expansion).
    python tests/gpu_fuzz.py [n_rounds=20] [seed0=0]      -> prints one line per round, exits 1 on the first mismatch"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_pkg, load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402

FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")


def scene_frame(rng, sp):
    """An organised range image: tilted ground plane with noise, random walls, dropouts, duplicates, -1 markers, label zeros."""
    N, H = sp.n_scan, sp.horizon_scan
    rows, cols = np.divmod(np.arange(N * H), H)
    az = cols / H * 2 * np.pi
    el = np.deg2rad(rng.uniform(1, 12) - rows / N * rng.uniform(20, 45))
    h0 = rng.uniform(1.2, 2.2); tilt = np.deg2rad(rng.uniform(-4, 4, 2))
    with np.errstate(all="ignore"):
        rg = np.where(el < -0.01, h0 / np.tan(-el), 200.0)
    wall = rng.uniform(3, 90, H // 8 + 1).repeat(8)[:H][cols] * rng.uniform(0.9, 1.1, N * H)
    use_wall = (rng.random(H // 16 + 1).repeat(16)[:H][cols] < rng.uniform(0.1, 0.6)) & (wall < rg)
    r = np.where(use_wall, wall, rg)
    x = r * np.cos(az); y = r * np.sin(az)
    z = np.where(use_wall, r * np.tan(el), -h0 + x * np.tan(tilt[0]) + y * np.tan(tilt[1]))
    z = z + rng.normal(0, rng.choice([0.0, 0.01, 0.03, 0.1]), N * H)
    keep = (r < 150) & (rng.random(N * H) > rng.uniform(0, 0.3))
    idx = np.nonzero(keep)[0]
    dup = rng.choice(idx, max(1, len(idx) // rng.integers(20, 400)))
    idx = rng.permutation(np.concatenate([idx, dup]))
    n = len(idx)
    jit = rng.normal(0, 0.02, (3, n))
    f = dict(x=x[idx] + jit[0], y=y[idx] + jit[1], z=z[idx] + jit[2],
             intensity=np.where(rng.random(n) < rng.choice([0.0, 0.01, 0.2]), -1.0, rng.random(n)),
             row=rows[idx], col=cols[idx], label=np.where(rng.random(n) < rng.choice([0.0, 0.05]), 0, rng.integers(-3, 4, n)))
    return {k: np.asarray(f[k]).astype(t) for k, t in cases.FIELD_TYPES}


def main():
    n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    pkg, synth, O = load_pkg(), load_synth(), load_oracle()
    gens = {}
    total = 0
    for rnd in range(n_rounds):
        rng = np.random.default_rng(1000 + seed0 + rnd)
        sensor = ("HDL_64E", "OS1_64", "HDL_32E")[rnd % 3]
        sp = O.sensor(sensor)
        frames = [scene_frame(rng, sp) for _ in range(5)]
        frames.append(cases.rand_frame(rng, sp.n_scan, sp.horizon_scan, int(rng.integers(0, 2 * sp.S)), spread=float(rng.uniform(5, 150)),
                                       zlo=float(rng.uniform(-12, -1)), zhi=float(rng.uniform(0, 40)), p_neg1=float(rng.uniform(0, 0.6))))
        frames.append(cases.hot_cell_frame(sp, seed=int(rng.integers(1 << 30)), n=int(rng.integers(1, sp.S + 1)), jitter=float(rng.uniform(0.01, 1.9))))
        frames.append(synth.make_frame(sensor, int(rng.integers(1 << 20))))
        offs = np.zeros(len(frames) + 1, np.int64); offs[1:] = np.cumsum([len(f["x"]) for f in frames])
        batch = {k: np.concatenate([f[k] for f in frames]) for k in FIELDS}; batch["offsets"] = offs
        max_pts = int(np.diff(offs).max()) + 64
        key = (sensor, max_pts > 2 * sp.S)
        if key not in gens:
            gens[key] = pkg.BevGen(sensor, device=0, max_frames_per_batch=3, max_points_per_frame=max(max_pts, 2 * sp.S + 64))
        g = gens[key]
        ref = O.frames(sp, offs, *[batch[k] for k in FIELDS], n_threads=os.cpu_count() or 1)
        out = g.process_host(batch)
        cin = {k: batch[k] for k in ("x", "y", "z")}; cin["meta"] = pkg.pack_meta(g.params, batch["row"], batch["col"], batch["intensity"], batch["label"]); cin["offsets"] = offs
        exp = g.compact_to_reference_layout(g.process_host_compact(cin), batch)
        bad = []
        for name, got in (("host", out), ("compact", exp)):
            for k in ("owner", "label", "single", "multi"):
                if not np.array_equal(got[k], ref[k]):
                    fr = [i for i in range(len(frames)) if not np.array_equal(got[k][i], ref[k][i])]
                    bad.append("%s.%s frames %s" % (name, k, fr))
        total += len(frames)
        print("round %d seed %d %s: %d frames, %d points, ground slots %d  %s" % (rnd, 1000 + seed0 + rnd, sensor, len(frames), int(offs[-1]),
              int(((ref["label"] == 0) & (ref["owner"] > 0)).sum()), "OK" if not bad else "MISMATCH " + "; ".join(bad)), flush=True)
        if bad:
            sys.exit(1)
    print("fuzz ok: %d frames bit-exact through both staging formats" % total)


if __name__ == "__main__":
    main()
