"""Generates the committed golden vectors from the CPU oracle (seeded, deterministic).

    python tests/golden/make_golden.py

The reference itself cannot run in this image (Python 2, h5py/pywrapfst/magphase absent) and
ships no fixtures, so the vectors are produced by oracle/snickery_oracle.py, whose k-NN stage IS
the reference's engine (scipy cKDTree with the reference's kwargs).  Inputs are stored next to
the expected outputs so the tests do not depend on the random generator's bit stream.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import epoch_config, halfphone_config  # noqa: E402
from oracle import snickery_oracle as O  # noqa: E402
from snickery_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_epoch():
    db = syn.make_epoch_db(n_units=1500, seed=1235)
    cfg = epoch_config()
    s = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    s.get_tree_for_greedy_search()
    targets = syn.make_targets(db["F"], 3, 121, seed=11)   # 121 frames: remainder 1 is cut at m=6
    out = {"F": db["F"], "Jc": db["Jc"]}
    for i, x in enumerate(targets):
        uf = O.weight(x, s.target_weight_vector)
        p_tree, d_tree = s.greedy_joint_search(uf, return_dists=True)
        p_brute = s.greedy_joint_search(uf, engine="brute")
        assert p_tree == p_brute
        out["targets_%d" % i] = uf
        out["path_%d" % i] = np.array(p_tree, dtype=np.int64)
        out["dist_%d" % i] = d_tree
        if i == 0:
            out["tscores_0"] = s.get_target_scores_per_stream(s.window_targets(uf), p_tree)
            out["jscores_0"] = s.get_join_scores_per_stream(p_tree)
    # identity: un-windowed consecutive DB frames from start_state (synth_simple.py:909-928 for m>1)
    start, n, m = 400, 12, cfg["multiepoch"]
    tf = s.train_unit_features[start:start + m * n]
    p = s.greedy_joint_search(tf, start_state=start)
    assert p == list(range(start, start + m * n, m))
    out["identity_start"] = np.int64(start)
    out["identity_path"] = np.array(p, dtype=np.int64)
    # multiepoch = 1 variant
    cfg1 = epoch_config(multiepoch=1)
    s1 = O.OracleSynthesiser(cfg1, db["F"], db["Jc"])
    s1.get_tree_for_greedy_search()
    uf = O.weight(targets[0][:40], s1.target_weight_vector)
    out["m1_targets"] = uf
    out["m1_path"] = np.array(s1.greedy_joint_search(uf), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "epoch_greedy.npz"), **out)


def make_halfphone():
    db = syn.make_halfphone_db(n_units=1200, seed=1237)
    cfg = halfphone_config(n_candidates=12)
    s = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    s.build_acoustic_tree()
    out = {"F": db["F"], "Jc": db["Jc"], "phones": db["phones"]}
    x = syn.make_targets(db["F"], 1, 30, seed=5)[0]
    uf = O.weight(x, s.target_weight_vector)
    cand, dist = s.preselect_units_acoustic(uf)
    out["targets"] = uf
    out["knn_idx"] = cand.astype(np.int64)
    out["knn_dist"] = dist
    for name, arith in (("f64", "f64"), ("fst32", "openfst32")):
        p, c = s.viterbi_search(cand, dist, arithmetic=arith, return_cost=True)
        out["vit_path_" + name] = np.array(p, dtype=np.int64)
        out["vit_cost_" + name] = np.float64(c)
    # quinphone-like lattice: duplicates, -1 padding, inadmissible ids 0 and N-1
    tphones = db["phones"][100:130]
    cq = syn.quinphone_like_candidates(db["phones"], tphones, 12, seed=9)
    cq[3, 0] = 0
    cq[4, 1] = db["F"].shape[0] - 1
    cq[5, :] = np.where(np.arange(12) < 2, cq[5, :], -1)
    dq = s.candidate_distances(cq, uf)
    out["q_cand"] = cq
    out["q_dist"] = dq
    p, c = s.viterbi_search(cq, dq, return_cost=True)
    p2, c2 = O.viterbi_search_numpy(s, cq, dq, return_cost=True)
    assert p == p2 and abs(c - c2) < 1e-9
    out["q_path"] = np.array(p, dtype=np.int64)
    out["q_cost"] = np.float64(c)
    tc, jc, tot = s.path_costs(cq, dq, p)
    out["q_tcost"], out["q_jcost"] = np.float64(tc), np.float64(jc)
    # join tile golden
    out["tile_0"] = np.array([[s.join_cost_cache(cq[:2]).get((int(a), int(b)), np.inf) for b in cq[1]] for a in cq[0]])
    # tiny lattice checked by exhaustive enumeration
    cs, ds = cand[:5, :4], dist[:5, :4]
    pe, ce = s.viterbi_exhaustive(cs, ds)
    pv, cv = s.viterbi_search(cs, ds, return_cost=True)
    assert pe == pv and abs(ce - cv) < 1e-9
    out["tiny_path"] = np.array(pe, dtype=np.int64)
    out["tiny_cost"] = np.float64(ce)
    np.savez_compressed(os.path.join(HERE, "halfphone_viterbi.npz"), **out)


if __name__ == "__main__":
    make_epoch()
    make_halfphone()
    for f in ("epoch_greedy.npz", "halfphone_viterbi.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
