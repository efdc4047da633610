import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _load_pkg import load_pkg, load_synth, load_oracle  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def synth():
    return load_synth()


@pytest.fixture(scope="session")
def O():
    return load_oracle()


FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")


def oracle_batch(O, sensor, batch, **kw):
    sp = O.sensor(sensor) if isinstance(sensor, str) else sensor
    return O.frames(sp, batch["offsets"], *[batch[k] for k in FIELDS], **kw)


def assert_same(out, ref, what=""):
    for k in ("owner", "label", "single", "multi"):
        a, b = np.asarray(out[k]), np.asarray(ref[k])
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s %s: %d mismatches, first at %s: got %s want %s" %
                                 (what, k, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def cat_frames(frames):
    offs = np.zeros(len(frames) + 1, np.int64)
    offs[1:] = np.cumsum([len(f["x"]) for f in frames])
    b = {k: np.concatenate([np.asarray(f[k]) for f in frames]) if frames else np.zeros(0) for k in FIELDS}
    b["offsets"] = offs
    return b
