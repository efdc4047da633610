"""Generates tests/golden/bev_golden.json (+ bev_golden_small.npz) from the REFERENCE'S OWN SOURCE: every vector below is
what /root/reference/BatchMultiBevGen.cpp computes when compiled unmodified against oracle/stub (oracle/_ref/
libbevgen_ref.so, recipe oracle/Makefile).  Run in the build container (where /root/reference exists):

    python tests/golden/make_bev_golden.py

The inputs are regenerated from seeds at test time (tests/cases.py, synth.py); the fixture stores, per frame, the sha256 of
the ordered cloud / owner / label / single / multi the reference produced, plus full arrays for a few small frames so a
mismatch can be located.  The GPU box has no /root/reference: there the CUDA path is compared with these vectors."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402

FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")


frame_list = cases.golden_frame_list


digest = cases.digest


def main():
    synth, O = load_synth(), load_oracle()
    assert O.ref_bevgen_lib() is not None, "build oracle/_ref first (make -C oracle ref)"
    meta = {"generator": "oracle/_ref/libbevgen_ref.so = /root/reference/BatchMultiBevGen.cpp + src/Utility.cpp @ d94040e, float overloads",
            "overloads": O.ref_math_overloads(), "frames": {}}
    small = {}
    for cid, sensor, f in frame_list(synth, O):
        n = len(f["x"])
        r = O.ref_frame(sensor, *[f[k] for k in FIELDS], t=np.arange(1, n + 1, dtype=np.uint32))
        rd = O.ref_frame(sensor, *[f[k] for k in FIELDS], double_libm=True)
        meta["frames"][cid] = dict(sensor=sensor, n=n, owner=digest(r["t"]), label=digest(r["label"]), single=digest(r["single"]),
                                   multi=digest(r["multi"]), x=digest(r["x"]), z=digest(r["z"]), label_double_libm=digest(rd["label"]),
                                   n_ground=int((r["label"] == 0).sum()), n_occupied=int((r["multi"] != 0).sum()))
        if sensor == "HDL_32E" and (cid.startswith("synth") or cid in ("boundary", "borderline/0", "hot/HDL_32E")):
            for k in ("label", "single", "multi"):
                small[cid + ":" + k] = r[k]
    for K, seed, step in cases.GOLDEN_LABEL_SETS:
        xyz = synth.make_poses(K, seed=seed, step=step)
        mi, lab = O.ref_select_and_label(xyz)
        nz = np.argwhere(lab != 0)
        meta.setdefault("labels", {})["K%d_s%d" % (K, seed)] = dict(K=K, seed=seed, step=step, M=int(len(mi)), major=digest(mi), labels=digest(lab))
        if K <= 1500:
            small["labels/K%d:major" % K] = mi
            small["labels/K%d:nz_index" % K] = nz.astype(np.int32)
            small["labels/K%d:nz_value" % K] = lab[lab != 0]
    d = os.path.dirname(os.path.abspath(__file__))
    json.dump(meta, open(os.path.join(d, "bev_golden.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(d, "bev_golden_small.npz"), **small)
    print("wrote %d frame digests, %d label sets, %d arrays" % (len(meta["frames"]), len(meta["labels"]), len(small)))


if __name__ == "__main__":
    main()
