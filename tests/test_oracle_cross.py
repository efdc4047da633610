"""The C oracle against the independent pure-Python restatement (oracle/bevgen_oracle_py.py), bit for bit,
on small sensor shapes and seeded random frames with every quirk switched on."""
import numpy as np
import pytest

from conftest import FIELDS


def rand_frame(rng, N, H, n, spread, p_neg1):
    t = (np.float32,) * 4 + (np.uint16, np.uint16, np.int16)
    v = dict(x=rng.uniform(-spread, spread, n), y=rng.uniform(-spread, spread, n), z=rng.uniform(-3, 8, n),
             intensity=np.where(rng.random(n) < p_neg1, -1.0, rng.random(n)), row=rng.integers(0, N + 1, n),
             col=rng.integers(0, H + 2, n), label=rng.integers(-2, 3, n))
    return {k: np.asarray(a).astype(tt) for (k, a), tt in zip(v.items(), t)}


def structured_frame(rng, N, H):
    rows, cols = np.divmod(np.arange(N * H), H)
    rad = 3.0 + (N - 1 - rows) * 1.2 + rng.normal(0, 0.05, N * H)
    a = 2 * np.pi * cols / H
    z = np.where(rng.random(N * H) < 0.15, rng.uniform(-1, 4, N * H), -1.7 + rng.normal(0, 0.03, N * H))
    keep = rng.random(N * H) > 0.1
    f = dict(x=(rad * np.cos(a))[keep], y=(rad * np.sin(a))[keep], z=z[keep],
             intensity=np.where(rng.random(keep.sum()) < 0.05, -1.0, 0.5), row=rows[keep], col=cols[keep], label=np.full(keep.sum(), -2))
    t = (np.float32,) * 4 + (np.uint16, np.uint16, np.int16)
    return {k: np.asarray(v).astype(tt) for (k, v), tt in zip(f.items(), t)}


@pytest.mark.parametrize("seed", range(4))
def test_c_oracle_equals_python_restatement(O, seed):
    import bevgen_oracle_py as P
    rng = np.random.default_rng(seed)
    N, H, G, hr = [(12, 40, 6, 0.5), (16, 33, 9, 0.25), (10, 64, 7, 1.0), (20, 21, 15, 0.5)][seed]
    sp = O.Sensor(); sp.n_scan, sp.horizon_scan, sp.ground_upper_scan, sp.height_res = N, H, G, hr
    for f in (rand_frame(rng, N, H, N * H * 2, 40.0, 0.2), structured_frame(rng, N, H), rand_frame(rng, N, H, 50, 150.0, 0.0)):
        oc = O.order(sp, *[f[k] for k in FIELDS])
        po = P.ordered_cloud(N, H, *[f[k] for k in FIELDS])
        for k in ("x", "y", "z", "intensity", "label", "owner"):
            assert np.array_equal(oc[k], po[k]), k
        lab, gm1, gmf, avg = O.mark_ground(sp, oc)
        plab, pgm1, pgmf, pavg = P.mark_ground(N, H, G, po)
        assert np.array_equal(gm1, pgm1) and np.array_equal(gmf, pgmf)
        assert np.array_equal(avg.view(np.uint32), pavg.view(np.uint32))
        assert np.array_equal(lab, plab)
        ps, pm = P.bevs(N, H, hr, po, plab)
        assert np.array_equal(O.single_bev(sp, oc, lab), ps)
        assert np.array_equal(O.multi_bev(sp, oc, lab), pm)
        full = O.frame(sp, *[f[k] for k in FIELDS])
        assert np.array_equal(full["label"], lab) and np.array_equal(full["single"], ps) and np.array_equal(full["multi"], pm)


def test_threads_do_not_change_results(O, synth):
    from conftest import oracle_batch
    b = synth.make_batch("HDL_32E", 5)
    a1 = oracle_batch(O, "HDL_32E", b, n_threads=1); a4 = oracle_batch(O, "HDL_32E", b, n_threads=4)
    for k in a1:
        assert np.array_equal(a1[k], a4[k])


def test_labels_python_vs_c(O, synth):
    import bevgen_oracle_py as P
    xyz = synth.make_poses(120, seed=3)
    mi, _ = O.select_major(xyz)
    assert np.array_equal(mi, P.select_major(xyz))
    lab, _, _ = O.labels(xyz, mi)
    assert np.array_equal(lab, P.labels(xyz, mi))
