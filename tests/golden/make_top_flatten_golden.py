"""Generates tests/golden/top_flatten_golden.npz from the REFERENCE'S OWN SOURCE: the flattened clouds that
/root/reference/TopPartRegistration.cpp's extractTopAndFlatten (:79-141) returns when the file is compiled unmodified against
oracle/stub (oracle/_ref/libtoppart_ref.so, recipe oracle/Makefile).  Run in the build container (where /root/reference exists):

    python tests/golden/make_top_flatten_golden.py

Inputs are regenerated from seeds at test time (tests/cases.py: top_flatten_cases); the fixture stores the sha256 of the output x / y arrays
of every case and the arrays themselves for the small ones (heights are distinct in all of them, so the order std::sort leaves is defined).  The GPU box has no
/root/reference: there bevgen_top_flatten is compared with these vectors (tests/test_golden_vectors.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _load_pkg import load_synth, load_oracle  # noqa: E402
import cases  # noqa: E402


def main():
    synth, O = load_synth(), load_oracle()
    out = {}
    for name, x, y, z, lab in cases.top_flatten_cases(O, synth):
        r = O.ref_top_flatten(x, y, z, lab)
        assert r is not None, "build oracle/_ref first (make -C oracle ref)"
        key = name.replace(" ", "_")
        out[key + ":n"] = np.int64(len(r[0]))
        out[key + ":sha256"] = np.array(cases.digest(np.concatenate([r[0], r[1]])))
        if len(r[0]) <= 2000:                  # full arrays for the small cases, so that a mismatch can be located
            out[key + ":x"], out[key + ":y"] = r
        print("%-16s %7d points in -> %6d out" % (name, len(x), len(r[0])))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "top_flatten_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
