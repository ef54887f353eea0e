"""Soundness stress of the exactness certificate: on large databases the tensor-core engine (certified or
re-searched) must agree with the exact-arithmetic SIMT engine on every query, up to float64 ties."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from conftest import halfphone_config  # noqa: E402
from snickery_b200 import Synthesiser, engine, synthetic as syn  # noqa: E402


def compare(db, q, k, space, label):
    db.set_engine(engine.ENGINE_TC)
    db.counters(reset=True)
    d1, i1 = db.knn(q, k, space)
    c = db.counters()
    db.set_engine(engine.ENGINE_SIMT)
    d2, i2 = db.knn(q, k, space)
    bad = i1 != i2
    nbad = int(bad.sum())
    worst = 0.0
    if nbad:
        rel = np.abs(d1 - d2)[bad] / np.maximum(d2[bad], 1e-300)
        worst = float(rel.max())
    print("%-34s queries=%6d k=%2d  index mismatches=%d (max rel distance gap %.2e)  recertified=%d" %
          (label, q.shape[0], k, nbad, worst, c["recertified"]), flush=True)
    assert worst <= 1e-6, "engines disagree beyond a tie"
    np.testing.assert_allclose(d1, d2, rtol=1e-9)


def main():
    cfg = bench.workload_config()
    dbe = bench.make_database(700000)
    g = Synthesiser(cfg, dbe["F"], dbe["Jc"])
    rng = np.random.default_rng(7)
    wt, wj = g.target_weight_vector, g.join_weight_vector
    Fw = dbe["F"].astype(np.float64) * wt
    Jw = dbe["Jc"].astype(np.float64) * wj
    for noise, nq in ((0.3, 20000), (0.02, 8000), (1.5, 4000)):
        rows = rng.integers(0, 699000, nq)
        q = np.hstack([Jw[rows]] + [Fw[rows + j] for j in range(6)])
        q = q + noise * rng.standard_normal(q.shape) * np.concatenate([wj, np.tile(wt, 6)])
        compare(g.db, q, 1, engine.SPACE_JOINT, "joint 517-d, noise %.2f" % noise)
    rows = rng.integers(0, 699000, 3000)
    q = np.hstack([Jw[rows]] + [Fw[rows + j] for j in range(6)]) + 0.1 * rng.standard_normal((3000, 517)) * 0.1
    compare(g.db, q, 20, engine.SPACE_JOINT, "joint 517-d, k=20")
    g.db.close()
    hp = syn.make_halfphone_db(n_units=90000, seed=1237)
    g3 = Synthesiser(halfphone_config(n_candidates=50), hp["F"], hp["Jc"])
    wt3 = g3.target_weight_vector
    for noise, nq in ((0.3, 20000), (0.02, 5000)):
        rows = rng.integers(0, 90000, nq)
        q = hp["F"][rows].astype(np.float64) * wt3 + noise * rng.standard_normal((nq, 184)) * wt3
        compare(g3.db, q, 50, engine.SPACE_TARGET, "half-phone 184-d, noise %.2f" % noise)
        compare(g3.db, q[:4000], 1, engine.SPACE_TARGET, "half-phone 184-d, noise %.2f" % noise)
    print("stress ok")


if __name__ == "__main__":
    main()
