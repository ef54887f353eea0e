"""Throughput of the greedy search for other multiepoch values (table-driven kernel path)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from conftest import epoch_config  # noqa: E402
from snickery_b200 import Synthesiser, engine  # noqa: E402

db = bench.make_database(700000)
for m in (1, 3, 4, 6):
    cfg = epoch_config(multiepoch=m, tsw=(0.5, 0.5))
    g = Synthesiser(cfg, db["F"], db["Jc"])
    B, T = 1024, 648 // 6 * m           # 108 steps each
    cat = bench.make_batch(db["F"], g.target_weight_vector, B, T, seed=m)
    lens = np.full(B, T, dtype=np.int64)
    g.db.greedy_batch_cat(cat, lens)
    g.db.profile_enable(True)
    t0 = time.perf_counter()
    g.db.greedy_batch_cat(cat, lens)
    dt = time.perf_counter() - t0
    p = g.db.profile_read(engine.PROF_KNN)
    c = g.db.counters()
    print("m=%d D=%d: %.1f ms/batch, %.2f M frames/s, GEMM %.0f TFLOP/s (%.3f ms/launch), recert %d/%d" %
          (m, 151 + 61 * m, dt * 1e3, B * T / dt / 1e6, p["work"] / p["ms"] / 1e9, p["ms"] / max(p["launches"], 1),
           c["recertified"], c["queries"]), flush=True)
    g.db.close()
