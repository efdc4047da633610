#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it loads the oracle and oracle/_ref).  The CPU half of tests/gpu_fuzz.py: the SAME seeded frames (scene
frames with random ground planes / walls / dropouts / -1 markers, random unstructured frames, hot-cell frames, synthetic keyframes;
all three sensors) through the reference's own source (oracle/_ref/libbevgen_ref.so = BatchMultiBevGen.cpp compiled unmodified)
and through the oracle: ordered cloud, owner, ground_mat, labels, both BEVs, .bin bytes, CSV text - for both overload sets of the
unqualified atan2 / sqrt; HDL_64E frames also through BatchCloudManip.cpp's own pipeline (labels + the 201 x 201 float map).  gpu_fuzz.py holds CUDA to the oracle on these frames, this holds the oracle to the reference.
    python tests/cpu_fuzz_ref.py [n_rounds=20] [seed0=0]      -> one line per round, exits 1 on the first mismatch
Needs oracle/_ref (built where /root/reference exists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402
from gpu_fuzz import scene_frame  # noqa: E402
from test_reference_source_pin import check_frame  # noqa: E402

FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")


def main():
    n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    synth, O = load_synth(), load_oracle()
    if O.ref_bevgen_lib() is None:
        sys.exit("oracle/_ref/libbevgen_ref.so not built")
    total = 0
    import tempfile
    tmp = tempfile.mkdtemp()
    for rnd in range(n_rounds):
        rng = np.random.default_rng(1000 + seed0 + rnd)                 # the frame recipe of gpu_fuzz.py, draw for draw
        sensor = ("HDL_64E", "OS1_64", "HDL_32E")[rnd % 3]
        sp = O.sensor(sensor)
        frames = [scene_frame(rng, sp) for _ in range(5)]
        frames.append(cases.rand_frame(rng, sp.n_scan, sp.horizon_scan, int(rng.integers(0, 2 * sp.S)), spread=float(rng.uniform(5, 150)),
                                       zlo=float(rng.uniform(-12, -1)), zhi=float(rng.uniform(0, 40)), p_neg1=float(rng.uniform(0, 0.6))))
        frames.append(cases.hot_cell_frame(sp, seed=int(rng.integers(1 << 30)), n=int(rng.integers(1, sp.S + 1)), jitter=float(rng.uniform(0.01, 1.9))))
        frames.append(synth.make_frame(sensor, int(rng.integers(1 << 20))))
        ground = differ = 0
        try:
            for i, f in enumerate(frames):
                _, o = check_frame(O, sensor, f, what="round %d frame %d" % (rnd, i))
                _, od = check_frame(O, sensor, f, double_libm=True, what="round %d frame %d (double libm)" % (rnd, i))
                ground += int(((o["label"] == 0) & (o["owner"] > 0)).sum())
                in_range = bool((f["row"] < sp.n_scan).all() and (f["col"] < sp.horizon_scan).all())   # its getOrderedCloud (:47-63) has no bounds test: out of range = a wild write
                if sensor == "HDL_64E" and in_range:      # BatchCloudManip.cpp's own order / ground / saveAsMat with the label filter (HDL-64E constants), SURVEY 8(f)-3
                    blab, bm = O.ref_bcm_frame(*[f[k] for k in FIELDS], out_prefix=os.path.join(tmp, "b"))
                    assert np.array_equal(blab, o["label"]), "round %d frame %d: batch_cloud_manip labels" % (rnd, i)
                    want = O.bvm(O.order(sp, *[f[k] for k in FIELDS]), o["label"])
                    assert np.array_equal(bm.view(np.uint32), want.view(np.uint32)), "round %d frame %d: batch_cloud_manip map" % (rnd, i)
                differ += int((o["label"] != od["label"]).sum())
        except AssertionError as e:
            print("round %d seed %d %s: MISMATCH %s" % (rnd, 1000 + seed0 + rnd, sensor, e), flush=True)
            sys.exit(1)
        total += len(frames)
        print("round %d seed %d %s: %d frames, %d points, ground slots %d, labels on which the two overload sets differ %d  OK" % (
            rnd, 1000 + seed0 + rnd, sensor, len(frames), sum(len(f["x"]) for f in frames), ground, differ), flush=True)
    print("fuzz ok: oracle == reference source on %d frames (both overload sets)" % total)


if __name__ == "__main__":
    main()
