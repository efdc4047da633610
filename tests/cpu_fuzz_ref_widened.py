#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it loads the oracle and oracle/_ref).  Seeded fuzz of the widened rows' oracle functions against the
reference's own source compiled unmodified (oracle/_ref): the three projections (extractPointCloud of the MulRan / Oxford / KITTI
extractors run on scan files), extractTopAndFlatten (TopPartRegistration.cpp), saveAsMat (CloudManip.cpp) - random clouds of random
size and spread, KITTI scans with random ring counts / short rings / sign flips, both overload sets for the projections.
    python tests/cpu_fuzz_ref_widened.py [n_rounds=20] [seed0=0]      -> one line per round, exits 1 on the first mismatch"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402


def bits(a):
    a = np.asarray(a)
    return np.where(np.isnan(a), np.float32(np.nan), a).view(np.uint32) if a.dtype == np.float32 else a


def main():
    n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    synth, O = load_synth(), load_oracle()
    devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 2)          # the extractors' initDirectories runs `rm -r` on missing folders
    with tempfile.TemporaryDirectory() as d:
        for rnd in range(n_rounds):
            rng = np.random.default_rng(5000 + seed0 + rnd)
            n = int(rng.integers(1, 400_000)); spread = float(rng.choice([0.5, 8.0, 60.0, 150.0, 1e4]))
            x = rng.normal(0, spread, n).astype(np.float32); y = rng.normal(0, spread, n).astype(np.float32)
            z = rng.normal(-1, float(rng.choice([0.05, 3.0, 40.0])), n).astype(np.float32)
            k = rng.integers(0, n, min(n, 500))                                 # axis points, zeros of both signs, exact diagonals
            x[k[0::5]] = 0.0; y[k[1::5]] = -0.0; x[k[2::5]] = y[k[2::5]]; y[k[3::5]] = 0.0; x[k[4::5]] = -x[k[4::5]]
            inten = rng.random(n).astype(np.float32)
            what = []
            for dbl in (False, True):
                m = min(n, 64 * 1024)
                r = O.ref_extract_point_cloud("mulran", d, x[:m], y[:m], z[:m], inten[:m], double_libm=dbl)
                row, col = O.project_mulran(x[:m], y[:m], double_libm=dbl)
                assert np.array_equal(r["row"][:m], row) and np.array_equal(r["col"][:m], col), ("mulran", rnd, dbl)
                r = O.ref_extract_point_cloud("oxford", d, x, y, z, inten, double_libm=dbl)
                nx, nz, row, col = O.project_oxford(x, y, z, double_libm=dbl)
                assert np.array_equal(r["row"], row) and np.array_equal(r["col"], col), ("oxford", rnd, dbl)
                assert np.array_equal(bits(r["x"]), bits(nx)) and np.array_equal(bits(r["z"]), bits(nz)), ("oxford xz", rnd, dbl)
                n_rings = int(rng.integers(1, 80)) if dbl is False else n_rings
                shorts = tuple(int(v) for v in rng.integers(0, max(n_rings, 1), int(rng.integers(0, 5)))) if dbl is False else shorts
                kw = dict(n_rings=n_rings, short_rings=shorts, start_negative=bool(rnd & 1), jitter=bool(rnd & 2))
                kx, ky, kz = synth.make_kitti_scan(10_000 + seed0 + rnd, **kw)
                kx, ky, kz = kx[:64 * cases.KITTI_H], ky[:64 * cases.KITTI_H], kz[:64 * cases.KITTI_H]     # the extractor's read limit (:172)
                r = O.ref_extract_point_cloud("kitti", d, kx, ky, kz, np.zeros(len(kx), np.float32), double_libm=dbl)
                row, col = O.project_kitti(kx, ky, double_libm=dbl)
                want = cases.kitti_structured(kx, ky, kz, row, col)
                for key in want:
                    assert np.array_equal(bits(r[key]), bits(want[key])), ("kitti", key, rnd, dbl, kw)
            what.append("projections %d pts, kitti %d rings / %d pts" % (n, n_rings, len(kx)))
            # extractTopAndFlatten: distinct heights (std::sort leaves the order of equal ones open)
            tz = (rng.permutation(n).astype(np.float32) * np.float32(0.0007) - np.float32(4.0))
            lab = rng.integers(-2, 3, n).astype(np.int16)
            tx = np.clip(x, -130, 130); ty = np.clip(y, -130, 130)
            rx, ry = O.ref_top_flatten(tx, ty, tz, lab)
            ox, oy, oi = O.top_flatten(tx, ty, tz, lab)
            assert len(rx) == len(ox) and np.array_equal(bits(rx), bits(ox)) and np.array_equal(bits(ry), bits(oy)), ("top_flatten", rnd)
            what.append("top_flatten %d -> %d" % (n, len(ox)))
            # saveAsMat (config #5), CSV text included
            csv = os.path.join(d, "m.csv")
            g = O.ref_save_as_mat(tx, ty, z, csv)
            w = O.save_as_mat(tx, ty, z)
            assert np.array_equal(bits(g), bits(w)), ("save_as_mat", rnd)
            what.append("save_as_mat %d cells" % int((w > 0).sum()))
            print("round %d seed %d: %s  OK" % (rnd, 5000 + seed0 + rnd, "; ".join(what)), flush=True)
    print("fuzz ok: widened-row oracle functions == reference source on %d rounds" % n_rounds)


if __name__ == "__main__":
    main()
