"""Timing probe over the shapes BASELINE.json lists (not a test, not the bench): prints ms per call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from conftest import halfphone_config  # noqa: E402
from snickery_b200 import Synthesiser, engine, synthetic as syn  # noqa: E402


def timeit(fn, n=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    cfg = bench.workload_config()
    db = bench.make_database(int(os.environ.get("PROBE_UNITS", "700000")))
    g = Synthesiser(cfg, db["F"], db["Jc"])
    for B in (1, 3, 32, 128, 512, 1024):
        cat = bench.make_batch(db["F"], g.target_weight_vector, B, 648, seed=B)
        lens = np.full(B, 648, dtype=np.int64)
        g.db.counters(reset=True)
        ms = timeit(lambda: g.db.greedy_batch_cat(cat, lens), n=2)
        c = g.db.counters()
        print("greedy  B=%4d: %8.2f ms/batch  %10.0f frames/s  (%.1f us/step, recert %d/%d)" %
              (B, ms, B * 648 / ms * 1e3, ms * 1e3 / 108, c["recertified"], c["queries"]), flush=True)
    hp = syn.make_halfphone_db(n_units=90000, seed=1237)
    for k in (1, 50):
        g3 = Synthesiser(halfphone_config(n_candidates=k), hp["F"], hp["Jc"])
        for nutt in (1, 64, 1024):
            uf = np.vstack(syn.make_targets(hp["F"], nutt, 80, seed=3)).astype(np.float64) * g3.target_weight_vector
            ms = timeit(lambda: g3.preselect_units_acoustic(uf), n=2)
            c = g3.db.counters(reset=True)
            print("acoustic k=%2d utts=%4d: %8.2f ms  %10.0f targets/s (recert %d/%d)" %
                  (k, nutt, ms, nutt * 80 / ms * 1e3, c["recertified"], c["queries"]), flush=True)


if __name__ == "__main__":
    main()
