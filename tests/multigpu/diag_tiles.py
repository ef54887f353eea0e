import numpy as np, sys
sys.path.insert(0, "tests")
from test_gpu_join_tc import *
hp = syn.make_halfphone_db(n_units=20000, seed=77)
cfg = halfphone_config(n_candidates=50, preselection="acoustic")
o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
g = Synthesiser(cfg, hp["F"], hp["Jc"])
for K in (50, 64, 8):
    rng = np.random.default_rng(K)
    n = hp["F"].shape[0]
    cands = [lattices(hp, rng, 24, K, n) for _ in range(3)]
    tiles = g.db.join_tiles(cands)
    fin_total, patched = g.db.join_stats()
    ref = np.concatenate([reference_tiles(o, c) for c in cands])
    fin = np.isfinite(ref)
    print("K", K, "inf pattern equal", np.array_equal(np.isinf(tiles), np.isinf(ref)), "finite", fin_total, int(fin.sum()), "patched", patched)
    pos = fin & (ref > 0) & np.isfinite(tiles)
    rel = np.abs(tiles[pos] - ref[pos]) / ref[pos]
    print("  rel err max %.3g  p99 %.3g  median %.3g" % (rel.max(), np.quantile(rel, 0.99), np.median(rel)))
    nn = None
# realistic lattices: acoustic preselection at 90k
hp = syn.make_halfphone_db(n_units=90000, seed=1237)
o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
g = Synthesiser(cfg, hp["F"], hp["Jc"])
utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(hp["F"], 4, 80, seed=31)]
cands = [g.preselect_units_acoustic(u)[0] for u in utts]
import os
for th in ("0.09375", "0.0625", "0.125"):
    os.environ["SNK_JOIN_THETA"] = th
    tiles = g.db.join_tiles(cands)
    fin_total, patched = g.db.join_stats()
    ref = np.concatenate([reference_tiles(o, c) for c in cands])
    pos = np.isfinite(ref) & (ref > 0)
    rel = np.abs(tiles[pos] - ref[pos]) / ref[pos]
    ratio = None
    print("theta", th, "finite", fin_total, "patched", patched, "frac %.4f" % (patched / fin_total), "rel err max %.3g p99 %.3g median %.3g" % (rel.max(), np.quantile(rel, 0.99), np.median(rel)))
# distribution of d^2 / (ne + ns)
c = cands[0]
e = o.unit_end_data[c[0]]; s = o.unit_start_data[c[1]]
d2 = ((e[:, None, :] - s[None, :, :]) ** 2).sum(axis=2)
nn = (e ** 2).sum(1)[:, None] + (s ** 2).sum(1)[None, :]
print("d2/(ne+ns) quantiles", np.quantile(d2 / nn, [0, 0.01, 0.1, 0.5, 0.9]))
