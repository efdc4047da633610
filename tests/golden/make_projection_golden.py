"""Generates tests/golden/projection_golden.json from the REFERENCE'S OWN SOURCE: what extractPointCloud of
/root/reference/{Mulran,Oxford,Kitti}PointCloudSelect.cpp returns when those files are compiled unmodified against oracle/stub
(oracle/_ref/lib{mulran,oxford,kitti}select_ref.so and the _dbl builds, recipe oracle/Makefile) and run on scan files written in the
datasets' layouts.  Run in the build container (where /root/reference exists):

    python tests/golden/make_projection_golden.py

Inputs are regenerated from seeds at test time (tests/cases.py: projection_cloud, KITTI_SCANS + synth.make_kitti_scan); the fixture
stores sha256 digests: row / col (and the negated x, z of Oxford) per point for MulRan / Oxford, the whole structured 64 x 2083 cloud
for KITTI, for both overload sets of the unqualified atan2 / sqrt / round.  The GPU box has no /root/reference: there bevgen_project
is compared with these vectors (tests/test_golden_vectors.py)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402

MULRAN_N = 64 * 1024          # MulranPointCloudSelect.cpp:110


def main():
    synth, O = load_synth(), load_oracle()
    x, y, z = cases.projection_cloud()
    inten = np.zeros(len(x), np.float32)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for dbl in (False, True):
            sfx = "_double_libm" if dbl else ""
            r = O.ref_extract_point_cloud("mulran", d, x[:MULRAN_N], y[:MULRAN_N], z[:MULRAN_N], inten[:MULRAN_N], double_libm=dbl)
            assert r is not None and len(r["x"]) == MULRAN_N, "build oracle/_ref first (make -C oracle ref)"
            out["mulran" + sfx] = {"n": MULRAN_N, "row": cases.digest(r["row"]), "col": cases.digest(r["col"])}
            r = O.ref_extract_point_cloud("oxford", d, x, y, z, inten, double_libm=dbl)
            assert len(r["x"]) == len(x)
            out["oxford" + sfx] = {"n": len(x), "row": cases.digest(r["row"]), "col": cases.digest(r["col"]),
                                   "x": cases.digest(r["x"]), "z": cases.digest(r["z"])}
            for seed, kw in cases.KITTI_SCANS:
                kx, ky, kz = synth.make_kitti_scan(seed, **kw)
                r = O.ref_extract_point_cloud("kitti", d, kx, ky, kz, np.zeros(len(kx), np.float32), double_libm=dbl)
                out["kitti_%d%s" % (seed, sfx)] = {"n": len(kx), "written_slots": int((r["label"] == -2).sum()),
                                                   "structured": cases.structured_digest(r)}
    for k, v in out.items():
        print(k, v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "projection_golden.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
