"""Generates tests/golden/labels_*.npz from the reference's own KD-tree (oracle/_ref/libnanoflann_ref.so, built from
/root/reference/include/nanoflann.hpp in place).  Run in the build container:  python tests/golden/make_label_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_oracle, load_synth  # noqa: E402
from test_oracle_labels_ref import _ref_pipeline  # noqa: E402

O, synth = load_oracle(), load_synth()
assert O.ref_lib() is not None, "build oracle/_ref first (make -C oracle ref)"
here = os.path.dirname(os.path.abspath(__file__))
cases = {"labels_k100_s7": synth.make_poses(100, seed=7), "labels_k400_s21": synth.make_poses(400, seed=21),
         "labels_k60_line": np.stack([np.arange(60) * 7.3, np.zeros(60), np.sin(np.arange(60))], 1).astype(np.float32)}
for name, xyz in cases.items():
    mi, lab = _ref_pipeline(O, xyz)
    np.savez_compressed(os.path.join(here, name + ".npz"), xyz=xyz, major_idx=mi, labels=lab)
    print(name, "K", len(xyz), "M", len(mi))
