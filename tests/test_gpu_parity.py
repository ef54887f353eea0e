"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden vectors.  Tolerances are BASELINE.json's: unit sequences bit-exact except at cost ties
within 1e-6 relative, costs within 1e-5 relative."""
import numpy as np
import pytest

from conftest import epoch_config, halfphone_config
from oracle import snickery_oracle as O
from snickery_b200 import GpuKDTree, GpuStashableKDTree, Synthesiser, engine, synthetic as syn

pytestmark = pytest.mark.gpu

TIE_RTOL = 1e-6    # north_star: index mismatches allowed only at cost ties within 1e-6 relative
COST_RTOL = 1e-5   # north_star: target / join / path costs within 1e-5 relative

ENGINES = [engine.ENGINE_SIMT, engine.ENGINE_TC]


def assert_knn_matches(dist, idx, ref_dist, ref_idx):
    """Same neighbours except where distances tie within TIE_RTOL; distances agree to COST_RTOL."""
    np.testing.assert_allclose(dist, ref_dist, rtol=COST_RTOL, atol=1e-12)
    bad = idx != ref_idx
    if bad.any():
        rel = np.abs(dist - ref_dist)[bad] / np.maximum(ref_dist[bad], 1e-300)
        assert np.all(rel <= TIE_RTOL), "index mismatch that is not a tie"
        # a tie may permute neighbours: each row must still hold the same multiset up to ties
        for q in np.flatnonzero(bad.any(axis=1)):
            for j in np.flatnonzero(bad[q]):
                tied = np.abs(ref_dist[q] - ref_dist[q, j]) <= TIE_RTOL * max(ref_dist[q, j], 1e-300)
                assert idx[q, j] in ref_idx[q, tied] or tied[-1], "neighbour set differs beyond a tie"


def assert_greedy_path_ok(oracle, uf, path, dists, start_state=-1):
    """Every GPU step must be optimal (within a tie) given the GPU's own previous choice."""
    ref_path, ref_d = oracle.greedy_joint_search(uf, start_state=start_state, return_dists=True)
    if path == ref_path:
        np.testing.assert_allclose(dists, ref_d, rtol=COST_RTOL, atol=1e-12)
        return 0
    wt = oracle.window_targets(np.asarray(uf, dtype=np.float64))
    n = oracle.current_join_rep.shape[1]
    prev = np.zeros(n) if start_state < 0 else oracle.prev_join_rep[start_state]
    assert len(path) == len(ref_path)
    ndiff = 0
    for t, ix in enumerate(path):
        d_all = oracle.greedy_step_distances(prev, wt[t])
        best = d_all.min()
        assert d_all[ix] <= best * (1 + TIE_RTOL) + 1e-300, "step %d: chose %d (%.17g) over optimum %.17g" % (t, ix, d_all[ix], best)
        assert abs(dists[t] - d_all[ix]) <= COST_RTOL * max(d_all[ix], 1e-12)
        ndiff += int(ix != int(np.argmin(d_all)))
        prev = oracle.current_join_rep[ix]
    return ndiff


@pytest.fixture(scope="module")
def epoch_pair(golden_epoch):
    cfg = epoch_config()
    o = O.OracleSynthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    return o, g


@pytest.fixture(scope="module")
def hp_pair(golden_halfphone):
    cfg = halfphone_config(n_candidates=12)
    o = O.OracleSynthesiser(cfg, golden_halfphone["F"], golden_halfphone["Jc"])
    o.build_acoustic_tree()
    g = Synthesiser(cfg, golden_halfphone["F"], golden_halfphone["Jc"])
    return o, g


# ------------------------------------------------------------------------------------ k-NN
@pytest.mark.parametrize("eng", ENGINES)
def test_knn_golden_halfphone(hp_pair, golden_halfphone, eng):
    o, g = hp_pair
    g.db.set_engine(eng)
    cand, dist = g.preselect_units_acoustic(golden_halfphone["targets"])
    assert cand.dtype == np.int64 and dist.dtype == np.float64 and cand.shape == (30, 12)
    assert_knn_matches(dist, cand, golden_halfphone["knn_dist"], golden_halfphone["knn_idx"])
    assert np.all(np.diff(dist, axis=1) >= 0)


@pytest.mark.parametrize("eng", ENGINES)
@pytest.mark.parametrize("k", [1, 7, 50])
def test_knn_joint_space_vs_ckdtree(epoch_pair, k, eng):
    o, g = epoch_pair
    g.db.set_engine(eng)
    rng = np.random.default_rng(k)
    combined = o.combined_rep()
    q = combined[rng.integers(0, combined.shape[0], 37)] + 0.05 * rng.standard_normal((37, combined.shape[1]))
    rd, ri = o.joint_tree.query(q, k=k)
    d, i = g.joint_tree.query(q, k=k)
    assert d.shape == rd.shape and i.shape == ri.shape
    assert_knn_matches(d.reshape(37, -1), i.reshape(37, -1), rd.reshape(37, -1), ri.reshape(37, -1))


def test_kdtree_adapter_scipy_conventions():
    rng = np.random.default_rng(1)
    data = rng.standard_normal((500, 14)).astype(np.float32).astype(np.float64)
    import scipy.spatial
    ref = scipy.spatial.cKDTree(data, leafsize=100, balanced_tree=False)
    t = GpuKDTree(data, leafsize=100, balanced_tree=False)
    assert t.exact_storage and t.n == 500 and t.m == 14
    x = rng.standard_normal((9, 14))
    for k in (1, 3):
        rd, ri = ref.query(x, k=k, eps=0.0)
        d, i = t.query(x, k=k, eps=10.0)          # eps is honoured trivially by an exact answer
        assert d.shape == rd.shape and i.shape == ri.shape
        assert np.array_equal(i, ri)
        np.testing.assert_allclose(d, rd, rtol=1e-12)
    rd, ri = ref.query(x[0], k=1)
    d, i = t.query(x[0], k=1)
    assert np.isscalar(d) and i == ri and abs(d - rd) <= 1e-12 * rd
    rd, ri = ref.query(x[0].reshape(1, -1), k=1)   # the shape greedy_joint_search uses (synth_simple.py:488-490)
    d, i = t.query(x[0].reshape(1, -1), k=1)
    assert d.shape == rd.shape == (1,) and i[0] == ri[0]
    # k larger than n: scipy pads with (inf, n)
    small = GpuKDTree(data[:5])
    d, i = small.query(x[:2], k=8)
    rd, ri = scipy.spatial.cKDTree(data[:5]).query(x[:2], k=8)
    assert np.array_equal(i, ri) and np.array_equal(np.isinf(d), np.isinf(rd))
    # exact duplicates: lowest index first
    dup = GpuKDTree(np.vstack([data[:10], data[:10]]))
    d, i = dup.query(data[3], k=2)
    assert i.tolist() == [3, 13] and np.all(d == 0.0)
    # sklearn-style wrapper
    s = GpuStashableKDTree(data, leaf_size=100, metric="euclidean")
    d, i = s.query(x, k=1)
    assert d.shape == (9, 1) and i.shape == (9, 1)


# ------------------------------------------------------------------------------------ greedy
@pytest.mark.parametrize("eng", ENGINES)
def test_greedy_golden(epoch_pair, golden_epoch, eng):
    o, g = epoch_pair
    g.db.set_engine(eng)
    utts = [golden_epoch["targets_%d" % i] for i in range(3)]
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    for i in range(3):
        assert len(paths[i]) == 121 // 6
        if paths[i] == golden_epoch["path_%d" % i].tolist():
            np.testing.assert_allclose(dists[i], golden_epoch["dist_%d" % i], rtol=COST_RTOL)
        else:
            assert_greedy_path_ok(o, utts[i], paths[i], dists[i])
    # single-utterance call site (synth_simple.py:413)
    assert g.greedy_joint_search(utts[0]) == paths[0]


@pytest.mark.parametrize("eng", ENGINES)
def test_greedy_identity_known_answer(epoch_pair, golden_epoch, eng):
    """The reference's own assertion (synth_simple.py:909-928): DB frames in, identity path out."""
    o, g = epoch_pair
    g.db.set_engine(eng)
    start, m, n = int(golden_epoch["identity_start"]), 6, 12
    tf = o.train_unit_features[start:start + m * n]
    path, d = g.greedy_joint_search_batch([tf], [start], return_dists=True)
    assert path[0] == golden_epoch["identity_path"].tolist()
    assert np.all(d[0] == 0.0)


@pytest.mark.parametrize("eng", ENGINES)
def test_greedy_ragged_batch_and_multiepoch1(golden_epoch, eng):
    cfg = epoch_config(multiepoch=1)
    o = O.OracleSynthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    g.db.set_engine(eng)
    p = g.greedy_joint_search(golden_epoch["m1_targets"])
    assert p == golden_epoch["m1_path"].tolist()
    tg = syn.make_targets(golden_epoch["F"], 5, 33, seed=21)
    utts = [O.weight(x[: n], o.target_weight_vector) for x, n in zip(tg, (33, 1, 17, 33, 8))]
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    for u, p, d in zip(utts, paths, dists):
        assert len(p) == u.shape[0]
        assert_greedy_path_ok(o, u, p, d)


@pytest.mark.parametrize("eng", ENGINES)
@pytest.mark.parametrize("m", [1, 3, 4])
def test_greedy_halfphone_epoch_join_layout(golden_epoch, eng, m):
    """Epoch voices written by train_halfphone.py store two-frame join windows; greedy search splits them
    into prev / current halves (synth_halfphone.py:552-553,580-581,693-695)."""
    F = golden_epoch["F"]
    Jc1 = golden_epoch["Jc"]
    Jc = np.ascontiguousarray(np.hstack([Jc1, np.vstack([Jc1[1:], Jc1[-1:]])]))     # [N+1, 302]: frame u-1 | frame u
    cfg = dict(epoch_config(multiepoch=m), halfphone_epoch_join_layout=True)
    o = O.OracleSynthesiser(cfg, F, Jc)
    o.get_tree_for_greedy_search()
    assert o.prev_join_rep.shape[1] == 151 and o.current_join_rep.shape[0] == F.shape[0] - (m - 1)
    g = Synthesiser(cfg, F, Jc)
    g.db.set_engine(eng)
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(F, 3, 30, seed=41)]
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    for u, p, d in zip(utts, paths, dists):
        assert len(p) == 30 // m
        assert_greedy_path_ok(o, u, p, d)


def test_greedy_too_short_utterance_raises(epoch_pair, golden_epoch):
    o, g = epoch_pair
    with pytest.raises(ValueError):
        g.greedy_joint_search(golden_epoch["targets_0"][:5])


def test_reweighting_matches_fresh_oracle(golden_epoch):
    """balance_stream_weights.py:84-88: new weights, rebuilt 'tree', search again."""
    cfg = epoch_config()
    g = Synthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    new = {"target_stream_weights": [0.7, 0.3], "join_stream_weights": [0.4, 0.1, 0.1, 0.4], "join_cost_weight": 0.35}
    g.reconfigure_settings(new)
    cfg2 = dict(cfg, **new)
    o = O.OracleSynthesiser(cfg2, golden_epoch["F"], golden_epoch["Jc"])
    o.get_tree_for_greedy_search()
    x = syn.make_targets(golden_epoch["F"], 1, 60, seed=33)[0]
    uf = O.weight(x, o.target_weight_vector)
    assert np.array_equal(g.target_weight_vector, o.target_weight_vector)
    paths, dists = g.greedy_joint_search_batch([uf], return_dists=True)
    assert_greedy_path_ok(o, uf, paths[0], dists[0])


def test_per_stream_scores(epoch_pair, golden_epoch):
    o, g = epoch_pair
    uf = golden_epoch["targets_0"]
    p = golden_epoch["path_0"].tolist()
    ts, js = g.get_scores_per_stream(uf, p)
    np.testing.assert_allclose(js, golden_epoch["jscores_0"], rtol=1e-12, atol=1e-300)
    # target scores: sum over the m frames of the step (equals the reference for m = 1)
    wt = o.window_targets(uf)
    sq = (o.windowed_unit_features[p] - wt) ** 2
    want = np.zeros((len(p), 2))
    for j in range(6):
        want[:, 0] += sq[:, j * 61: j * 61 + 60].sum(axis=1)
        want[:, 1] += sq[:, j * 61 + 60]
    np.testing.assert_allclose(ts, want, rtol=1e-12)


# ------------------------------------------------------------------------------------ join + Viterbi
def test_candidate_distances_golden(hp_pair, golden_halfphone):
    o, g = hp_pair
    d = g.candidate_target_distances(golden_halfphone["q_cand"], golden_halfphone["targets"])
    np.testing.assert_allclose(d, golden_halfphone["q_dist"], rtol=1e-12)


def test_join_tiles_vs_oracle(hp_pair, golden_halfphone):
    o, g = hp_pair
    cand = golden_halfphone["q_cand"]
    tiles = g.db.join_tiles([cand])
    assert tiles.shape == (cand.shape[0] - 1, 12, 12) and tiles.dtype == np.float32
    cache = o.join_cost_cache(cand)
    for t in range(cand.shape[0] - 1):
        want = np.array([[cache.get((int(a), int(b)), np.inf) for b in cand[t + 1]] for a in cand[t]])
        assert np.array_equal(np.isinf(tiles[t]), np.isinf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(tiles[t][fin], want[fin], rtol=COST_RTOL, atol=0)
        assert np.all(tiles[t][fin & (want == 0)] == 0)         # natural joins stay exactly free
    np.testing.assert_allclose(tiles[0], golden_halfphone["tile_0"], rtol=COST_RTOL)


def check_viterbi(o, cand, dist, path, pcost, tcost, jcost):
    ref_path, ref_cost = O.viterbi_search_numpy(o, cand, dist, return_cost=True)
    if not ref_path:
        assert path == []
        return
    assert len(path) == len(ref_path)
    tc, jc, tot = o.path_costs(cand, dist, path)       # float64 cost of the GPU's path
    assert tot <= ref_cost * (1 + TIE_RTOL), "GPU path is worse than the optimum beyond a tie"
    if path != ref_path:
        assert abs(tot - ref_cost) <= TIE_RTOL * ref_cost
    assert abs(pcost - tot) <= COST_RTOL * tot
    assert abs(tcost - tc) <= COST_RTOL * max(tc, 1e-12) and abs(jcost - jc) <= COST_RTOL * max(jc, 1e-12) + 1e-12


def test_viterbi_golden(hp_pair, golden_halfphone):
    o, g = hp_pair
    gh = golden_halfphone
    paths, pc, tc, jc = g.viterbi_search_batch([gh["knn_idx"], gh["q_cand"]], [gh["knn_dist"], gh["q_dist"]],
                                               return_costs=True)
    assert paths[0] == gh["vit_path_f64"].tolist() or paths[0] == gh["vit_path_fst32"].tolist()
    assert abs(pc[0] - float(gh["vit_cost_f64"])) <= COST_RTOL * float(gh["vit_cost_f64"])
    assert paths[1] == gh["q_path"].tolist()
    assert abs(pc[1] - float(gh["q_cost"])) <= COST_RTOL * float(gh["q_cost"])
    assert abs(tc[1] - float(gh["q_tcost"])) <= COST_RTOL * float(gh["q_tcost"])
    assert abs(jc[1] - float(gh["q_jcost"])) <= COST_RTOL * float(gh["q_jcost"])
    tiny = g.viterbi_search(gh["knn_idx"][:5, :4], gh["knn_dist"][:5, :4])
    assert tiny == gh["tiny_path"].tolist()
    # the single-utterance call site (synth_halfphone.py:1625)
    assert g.viterbi_search(gh["knn_idx"], gh["knn_dist"]) == paths[0]


def test_viterbi_quirks(hp_pair):
    o, g = hp_pair
    n = o.unit_end_data.shape[0]
    d = np.ones((3, 2))
    assert g.viterbi_search(np.array([[5, 6]]), np.ones((1, 2))) == []
    assert g.viterbi_search(np.array([[0, 0], [5, 6], [7, 8]]), d) == []
    assert g.viterbi_search(np.array([[4, 5], [n - 1, n - 1], [7, 8]]), d) == []
    assert g.viterbi_search(np.array([[4, -1], [5, 5], [-1, 6]]), d) == [4, 5, 6]
    paths, pc, tc, jc = g.viterbi_search_batch([np.array([[4, 900], [5, 901], [6, 902]])], [np.zeros((3, 2))], True)
    assert pc[0] == 0.0 and paths[0] == [4, 5, 6]         # lowest column wins the exact tie


@pytest.mark.parametrize("K", [1, 5, 30, 50, 64])
def test_viterbi_random_lattices(hp_pair, K):
    o, g = hp_pair
    rng = np.random.default_rng(100 + K)
    n = o.unit_end_data.shape[0]
    cands, dists = [], []
    for b in range(6):
        T = int(rng.integers(2, 40))
        c = rng.integers(1, n - 1, size=(T, K))
        c[rng.random((T, K)) < 0.15] = -1
        if b % 2:
            runs = rng.integers(1, n - T - 2)
            c[:, 0] = np.arange(runs, runs + T)               # a natural run is always available
        if b == 3:
            c[rng.integers(0, T)] = -1                         # a fully padded frame blocks every path
        c[0, -1] = 0
        c[-1, -1] = n - 1
        cands.append(c)
        dists.append(o.candidate_distances(c, rng.standard_normal((T, 184))) if b % 3 else rng.random((T, K)))
    paths, pc, tc, jc = g.viterbi_search_batch(cands, dists, return_costs=True)
    for b in range(6):
        check_viterbi(o, cands[b], dists[b], paths[b], pc[b], tc[b], jc[b])


def test_viterbi_beam1_is_greedy_over_candidates(hp_pair, golden_halfphone):
    o, g = hp_pair
    cand, dist = golden_halfphone["knn_idx"], golden_halfphone["knn_dist"]
    path = g.viterbi_search_batch([cand], [dist], greedy=True)[0]
    # restate: keep only the cheapest reachable state after every step
    T, K = cand.shape
    cache = o.join_cost_cache(cand)
    cur = {j: 0.0 for j in range(K)}
    chosen = []
    for t in range(T - 1):
        nxt = {}
        for jb in range(K):
            best = min(((cur[ja] + dist[t, ja] + cache[(int(cand[t, ja]), int(cand[t + 1, jb]))], ja)
                        for ja in cur if (int(cand[t, ja]), int(cand[t + 1, jb])) in cache), default=None)
            if best is not None:
                nxt[jb] = best
        jb = min(nxt, key=lambda j: (nxt[j][0], j))
        chosen.append(nxt[jb][1])
        cur = {jb: nxt[jb][0]}
    chosen.append(list(cur)[0])
    assert path[1:] == [int(cand[t, j]) for t, j in enumerate(chosen)][1:]


# ------------------------------------------------------------------------------------ larger, tensor-core engine
def _joint_d2(o, Q):
    """float64 squared joint distances of queries Q to every searchable row (BLAS form; only used to
    audit 1e-6 optimality, its own error is ~1e-13 relative)."""
    C = np.hstack([o.prev_join_rep, o.windowed_unit_features])
    return (Q * Q).sum(1)[:, None] + (C * C).sum(1)[None, :] - 2.0 * Q @ C.T


@pytest.mark.parametrize("eng", [engine.ENGINE_TC, engine.ENGINE_SIMT])
def test_greedy_batched_medium_database(eng):
    db = syn.make_epoch_db(n_units=30000, seed=77)
    cfg = epoch_config(tsw=(0.5, 0.5))
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.combined_rep()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    g.db.set_engine(eng)
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(db["F"], 48, 50, seed=5)]
    utts[3] = utts[3][:13]
    utts[7] = o.train_unit_features[1000:1048]                       # natural run: zero-distance answers
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    c = g.db.counters()
    assert c["recertified"] <= 0.02 * c["queries"], c                # the certificate almost always holds
    n = o.current_join_rep.shape[1]
    for b in (0, 3, 7, 20, 47):
        wt = o.window_targets(utts[b])
        assert len(paths[b]) == wt.shape[0]
        prev = np.zeros(n)
        for t, ix in enumerate(paths[b]):
            q = np.concatenate([prev, wt[t]])[None, :]
            d_all = np.sqrt(((np.hstack([o.prev_join_rep, o.windowed_unit_features]) - q) ** 2).sum(1)) \
                if t < 2 else np.sqrt(np.maximum(_joint_d2(o, q)[0], 0))
            assert d_all[ix] <= d_all.min() * (1 + TIE_RTOL) + 1e-9
            if t < 2:
                assert abs(dists[b][t] - d_all[ix]) <= COST_RTOL * max(d_all[ix], 1e-12)
            prev = o.current_join_rep[ix]
    # natural run: once a step lands on the run's own row, the next join is free and the chain must stay on it
    for t in range(1, len(paths[7])):
        if paths[7][t - 1] == 1000 + 6 * (t - 1):
            assert paths[7][t] == 1000 + 6 * t and dists[7][t] == 0.0


@pytest.mark.parametrize("k", [1, 4, 50])
def test_knn_tensor_core_halfphone_medium(k):
    hp = syn.make_halfphone_db(n_units=20000, seed=91)
    cfg = halfphone_config(n_candidates=k)
    o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
    o.build_acoustic_tree()
    g = Synthesiser(cfg, hp["F"], hp["Jc"])
    g.db.set_engine(engine.ENGINE_TC)
    uf = O.weight(np.vstack(syn.make_targets(hp["F"], 3, 70, seed=8)), o.target_weight_vector)
    cand, dist = g.preselect_units_acoustic(uf)
    rc, rd = o.preselect_units_acoustic(uf)
    assert_knn_matches(np.asarray(dist).reshape(210, -1), np.asarray(cand).reshape(210, -1),
                       np.asarray(rd).reshape(210, -1), np.asarray(rc).reshape(210, -1))
    c = g.db.counters()
    assert c["recertified"] <= 0.05 * c["queries"], c


# ------------------------------------------------------------------------------------ label-driven preselection, truncation
def _halfphone_names(phones):
    """Synthetic internal quinphone labels 'll/l/c_X/r/rr' (const.py:5, label_manip.py:16-32)."""
    names = []
    for i, p in enumerate(phones):
        side = "_L" if i % 2 == 0 else "_R"
        l, r = phones[max(i - 1, 0)], phones[min(i + 1, len(phones) - 1)]
        names.append("p%d/p%d/p%d%s/p%d/p%d" % (phones[max(i - 2, 0)], l, p, side, r, phones[min(i + 2, len(phones) - 1)]))
    return names


def test_monophone_then_acoustic_and_quinphone_preselection(golden_halfphone):
    gh = golden_halfphone
    names = _halfphone_names(gh["phones"].tolist())
    names[7] = "px/px/rare_L/px/px"                         # a phone with a single unit: padding path
    cfg = halfphone_config(n_candidates=6, preselection="monophone_then_acoustic")
    o = O.OracleSynthesiser(cfg, gh["F"], gh["Jc"])
    o.build_phonetrees(names)
    g = Synthesiser(cfg, gh["F"], gh["Jc"], train_unit_names=names)
    tnames = [names[i] for i in (100, 101, 7, 300, 301, 302)]
    uf = gh["targets"][:6]
    rc, rd = o.preselect_units_monophone_then_acoustic(uf, tnames)
    c, d = g.preselect_units_monophone_then_acoustic(uf, tnames)
    assert np.array_equal(c, rc)
    np.testing.assert_allclose(d, rd, rtol=1e-9)
    assert c[2, 1] == -1 and d[2, 1] == 1e15                  # const.VERY_BIG_WEIGHT_VALUE padding
    # quinphone label lookup + GPU target distances
    cq, dq = g.preselect_units_quinphone(uf, tnames)
    rq, rdq = o.preselect_units_quinphone(uf, tnames, g.unit_index)
    assert np.array_equal(cq, rq)
    np.testing.assert_allclose(dq, rdq, rtol=1e-12)
    # and the lattice goes through Viterbi like any other
    path = g.viterbi_search(cq, dq)
    assert path == O.viterbi_search_numpy(o, rq, rdq)


def test_truncated_streams(golden_epoch):
    """truncate_target_streams / truncate_join_streams (synth_simple.py:136-139,968-992)."""
    cfg = dict(epoch_config(multiepoch=2), truncate_target_streams=[20, -1], truncate_join_streams=[30, 10, 0, 1])
    o = O.OracleSynthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    o.truncate_target_streams(cfg["truncate_target_streams"])
    o.truncate_join_streams(cfg["truncate_join_streams"])
    o.get_tree_for_greedy_search()
    assert o.joint_tree.m == 41 + 2 * 21
    g = Synthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    x = syn.make_targets(golden_epoch["F"], 1, 40, seed=77)[0]
    full = O.weight(x, O.per_coeff_weights(np.array(cfg["target_stream_weights"]) * 0.8, cfg["stream_list_target"],
                                           cfg["datadims_target"]))
    uf = full[:, o.target_truncation_vector]                   # what synth_utt passes on (synth_simple.py:392-397)
    ref, rd = o.greedy_joint_search(uf, return_dists=True)
    paths, dists = g.greedy_joint_search_batch([uf], return_dists=True)
    assert paths[0] == ref
    np.testing.assert_allclose(dists[0], rd, rtol=COST_RTOL)


def test_certificate_failure_paths_stay_exact(golden_epoch, golden_halfphone, monkeypatch):
    """Pretend every 3rd query failed its exactness certificate: the SIMT re-search (k-NN) and the
    end-of-batch repair pass (greedy, deferred certificates) must still return the oracle's answer."""
    monkeypatch.setenv("SNK_DEBUG_CERT_FAIL", "3")
    cfg = epoch_config()
    o = O.OracleSynthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, golden_epoch["F"], golden_epoch["Jc"])
    g.db.set_engine(engine.ENGINE_TC)
    utts = [golden_epoch["targets_%d" % i] for i in range(3)] + [golden_epoch["targets_0"][:37]]
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    for u, p, d in zip(utts, paths, dists):
        assert_greedy_path_ok(o, u, p, d)
    assert g.db.counters()["recertified"] >= 1
    cfg3 = halfphone_config(n_candidates=12)
    o3 = O.OracleSynthesiser(cfg3, golden_halfphone["F"], golden_halfphone["Jc"])
    o3.build_acoustic_tree()
    g3 = Synthesiser(cfg3, golden_halfphone["F"], golden_halfphone["Jc"])
    g3.db.set_engine(engine.ENGINE_TC)
    cand, dist = g3.preselect_units_acoustic(golden_halfphone["targets"])
    assert_knn_matches(dist, cand, golden_halfphone["knn_dist"], golden_halfphone["knn_idx"])
    assert g3.db.counters()["recertified"] >= 10


# ------------------------------------------------------------------------------------ BASELINE.json full size (configs[1])
def test_full_size_database_properties():
    """700k-unit database (IS2018_nick_simplified shape), size-independent properties:
    identity known answer, tensor-core engine == exact-arithmetic SIMT engine, spot float64 audit,
    idempotence under re-weighting with the same weights."""
    import bench
    cfg = bench.workload_config()
    db = bench.make_database(bench.DB_UNITS)
    g = Synthesiser(cfg, db["F"], db["Jc"])
    wt = g.target_weight_vector
    F64 = db["F"].astype(np.float64) * wt
    # (1) the reference's known answer at full size: consecutive database frames give the identity path
    starts = [5, 123456, 699000 - 6 * 40]
    utts = [F64[s:s + 6 * 40] for s in starts]
    paths, dists = g.greedy_joint_search_batch(utts, starts, return_dists=True)
    for s, p, d in zip(starts, paths, dists):
        assert p == list(range(s, s + 6 * 40, 6))
        assert np.all(d == 0.0)
    # (2) noisy targets: tensor-core path == SIMT path (independent arithmetic), including distances
    cat = bench.make_batch(db["F"], wt, 24, 60, seed=4242)
    lens = np.full(24, 60, dtype=np.int64)
    g.db.set_engine(engine.ENGINE_TC)
    p_tc, d_tc = g.db.greedy_batch_cat(cat, lens, return_dists=True)
    c = g.db.counters()
    g.db.set_engine(engine.ENGINE_SIMT)
    p_simt, d_simt = g.db.greedy_batch_cat(cat, lens, return_dists=True)
    assert c["recertified"] <= 2
    for a, b, da, dbb in zip(p_tc, p_simt, d_tc, d_simt):
        if a != b:   # only a tie within 1e-6 may separate the two engines
            t = next(i for i in range(len(a)) if a[i] != b[i])
            assert abs(da[t] - dbb[t]) <= TIE_RTOL * dbb[t]
        else:
            np.testing.assert_allclose(da, dbb, rtol=1e-12)
    # (3) float64 audit of the first two steps of one utterance against a numpy scan of all 699,995 rows
    Jw = db["Jc"].astype(np.float64) * g.join_weight_vector
    prev = np.zeros(151)
    for t in range(2):
        q = np.concatenate([prev, cat[t * 6:(t + 1) * 6].reshape(-1)])
        d2 = ((Jw[:699995] - q[:151]) ** 2).sum(1)
        for j in range(6):
            d2 += ((F64[j:699995 + j] - q[151 + 61 * j:151 + 61 * (j + 1)]) ** 2).sum(1)
        ix = p_tc[0][t]
        assert d2[ix] <= d2.min() * (1 + 2 * TIE_RTOL)
        assert abs(np.sqrt(d2[ix]) - d_tc[0][t]) <= COST_RTOL * d_tc[0][t]
        prev = Jw[ix + 6]
    # (4) re-weighting with the same weights changes nothing
    g.db.set_engine(engine.ENGINE_TC)
    g.reconfigure_settings({})
    assert g.db.greedy_batch_cat(cat, lens) == p_tc


def test_stream_weight_balancing_loop_matches_cpu_loop(golden_epoch):
    """Row N1: the RPROP balancing loop (balance_stream_weights.py:82-172) driven by the GPU engine follows
    the same weight trajectory as the same loop driven by the oracle."""
    from snickery_b200 import balance
    cfg = epoch_config(multiepoch=1)
    F, Jc = golden_epoch["F"], golden_epoch["Jc"]
    tune = [x.astype(np.float64) for x in syn.make_targets(F, 4, 40, seed=88)]
    g = Synthesiser(cfg, F, Jc)
    best_g, losses_g, hist_g = balance.balance_stream_weights(g, tune, max_epochs=5)

    o = O.OracleSynthesiser(cfg, F, Jc)

    def evaluate(jw, tw):
        o.set_join_weights(jw)
        o.set_target_weights(tw)
        o.get_tree_for_greedy_search()
        js, ts = [], []
        for u in tune:
            uf = O.weight(u, o.target_weight_vector)
            p = o.greedy_joint_search(uf)
            ts.append(o.get_target_scores_per_stream(uf, p))
            js.append(o.get_join_scores_per_stream(p))
        return np.vstack(js), np.vstack(ts)

    best_o, losses_o, hist_o = balance.rprop_balance(evaluate, 4, 2, max_epochs=5)
    assert len(losses_g) == len(losses_o)
    np.testing.assert_allclose(losses_g, losses_o, rtol=1e-9)
    np.testing.assert_allclose(np.array(hist_g), np.array(hist_o), rtol=1e-12)
    np.testing.assert_allclose(best_g, best_o, rtol=1e-12)


# ------------------------------------------------------------------------------------ degenerate shapes and error paths
def test_degenerate_shapes_and_errors():
    import snickery_b200
    rng = np.random.default_rng(0)
    F = rng.standard_normal((7, 5)).astype(np.float32)
    Jc = rng.standard_normal((8, 3)).astype(np.float32)
    db = snickery_b200.UnitDatabase(F, Jc, multiepoch=2)
    with pytest.raises(snickery_b200.EngineError, match="set_weights"):
        db.knn(np.zeros((1, 5)), 1)
    db.set_weights(np.ones(5), np.ones(3))
    # empty query sets / batches
    d, i = db.knn(np.zeros((0, 5)), 3)
    assert d.shape == (0, 3) and i.shape == (0, 3)
    assert db.greedy_batch([]) == []
    paths, pc, tc, jc = db.join_viterbi_batch([], [])
    assert paths == []
    # k larger than the database, tiny database, joint space of 6 rows
    d, i = db.knn(F[:2].astype(np.float64), 9)
    assert np.all(i[:, 7:] == 7) and np.all(np.isinf(d[:, 7:])) and i[0, 0] == 0 and i[1, 0] == 1
    q = np.concatenate([Jc[3].astype(np.float64), F[3].astype(np.float64), F[4].astype(np.float64)])[None, :]
    d, i = db.knn(q, 2, engine.SPACE_JOINT)
    assert i[0, 0] == 3 and d[0, 0] == 0.0
    # greedy: the m-frame remainder is cut, one frame short of m raises like segment_axis
    p = db.greedy_batch([F[:5].astype(np.float64), F[:2].astype(np.float64)])
    assert [len(x) for x in p] == [2, 1]
    with pytest.raises(ValueError):
        db.greedy_batch([F[:1].astype(np.float64)])
    # Viterbi: zero-length and single-frame utterances inside a batch give empty paths
    cand = [np.array([[1, 2], [2, 3], [3, 4]]), np.zeros((0, 2), dtype=np.int64), np.array([[1, 2]]), np.array([[1, 1], [2, 2]])]
    dist = [np.ones((3, 2)), np.zeros((0, 2)), np.ones((1, 2)), np.ones((2, 2))]
    paths, pc, tc, jc = db.join_viterbi_batch(cand, dist)
    assert paths[0] == [1, 2, 3] and paths[1] == [] and paths[2] == [] and paths[3] == [1, 2]
    assert pc[0] == 3.0 and jc[0] == 0.0 and np.isinf(pc[1]) and np.isinf(pc[2])
    # bad arguments are errors, not crashes
    with pytest.raises(ValueError):
        db.knn(np.zeros((2, 4)), 1)
    with pytest.raises(snickery_b200.EngineError):
        db.join_viterbi_batch([np.ones((2, 200), dtype=np.int64)], [np.ones((2, 200))])    # K beyond the kernel's limit
    with pytest.raises(snickery_b200.EngineError):
        snickery_b200.UnitDatabase(F, Jc, multiepoch=0)
    # a database shorter than the multiepoch window has no joint rows
    short = snickery_b200.UnitDatabase(F[:2], Jc[:3], multiepoch=4)
    short.set_weights(np.ones(5), np.ones(3))
    with pytest.raises(snickery_b200.EngineError, match="empty"):
        short.knn(np.zeros((1, 3 + 4 * 5)), 1, engine.SPACE_JOINT)
    db.close()
    db.close()   # idempotent


# ------------------------------------------------------------------------------------ row N2: MagPhase concatenation
def _magphase_voice(n_sent=5, width=1025, seed=3):
    rng = np.random.default_rng(seed)
    sentences, names, within, lens = {}, [], [], []
    for i in range(n_sent):
        n = int(rng.integers(12, 30))
        f0 = np.where(rng.random((n, 1)) < 0.6, 100.0 + 50.0 * rng.random((n, 1)), 0.0).astype(np.float32)
        sentences["utt%d" % i] = tuple(rng.standard_normal((n, width)).astype(np.float32) for _ in range(3)) + (f0,)
        names += ["utt%d" % i] * n
        within += list(range(n))
        lens.append(n)
    return sentences, names, np.array(within), lens


@pytest.mark.parametrize("m,overlap", [(1, 0), (6, 2), (6, 6), (3, 0), (2, 2)])
def test_magphase_epoch_concatenation(m, overlap):
    from snickery_b200 import FrameStore
    sentences, names, within, lens = _magphase_voice()
    o = O.OracleMagPhaseStore(sentences, names, within, multiepoch=m)
    order = list(sentences)
    offs = np.concatenate([[0], np.cumsum(lens)])
    cat = [np.vstack([sentences[k][j] for k in order]) for j in range(3)]
    f0i, vuv = zip(*[O.lin_interp_f0(sentences[k][3]) for k in order])
    sent_of = np.repeat(np.arange(len(order)), lens)
    fs = FrameStore(cat[0], cat[1], cat[2], np.vstack(f0i), np.vstack(vuv), unit_frame=offs[sent_of] + within,
                    sent_lo=offs[sent_of], sent_hi=offs[sent_of + 1])
    rng = np.random.default_rng(m * 10 + overlap)
    n_units = len(names)
    # units whose m-frame body stays inside their sentence (the reference slices past the end otherwise),
    # including first / last units so the zero-padded edge fragments are exercised
    ok = np.flatnonzero(within + m <= np.array(lens)[sent_of])
    path = [int(ok[0])] + rng.choice(ok, 25).tolist() + [int(ok[-1])]
    want = o.concatenate(path, overlap=overlap)
    got = fs.concatenate(path, multiepoch=m, overlap=overlap)
    for w, g in zip(want, got):
        assert g.shape == w.shape and g.dtype == np.float64
        assert np.array_equal(g, w)          # float64 sums of the same float32 x float64 products: bit exact
    fzero = rng.random((len(path) * m, 1))
    assert fs.concatenate(path, multiepoch=m, overlap=overlap, fzero=fzero)[3] is not None
    with pytest.raises(AssertionError):
        fs.concatenate(path, multiepoch=m, overlap=1)


# ------------------------------------------------------------------------------------ shape fuzz: every kernel path
SHAPES = [
    # (Dt, Dj, m, N, k)   -- chosen to hit: embedded norms on/off, frame slab with 1-2 column blocks, plain multi-load
    (61, 151, 6, 3000, 1),     # static schedule 16
    (61, 151, 2, 2000, 3),     # table-driven, slab, embedded norms
    (64, 151, 3, 2500, 1),     # Dt = 64: no spare columns -> norm staging path
    (61, 64, 4, 2200, 2),      # Djq = 64: join part without spare columns -> norm staging path
    (70, 30, 2, 1800, 1),      # two column blocks per frame, slab serves both
    (130, 20, 3, 1500, 4),     # three column blocks per frame: 2 + 9 K-blocks > 10 -> SIMT engine takes over
    (20, 10, 9, 4000, 1),      # m = 9: largest window a slab serves
    (20, 10, 10, 4000, 1),     # m = 10: per-frame loads, table-driven
    (184, 8, 1, 5000, 7),      # static schedule 2 (target space queried below as well), store mode (tiny database)
    (33, 200, 1, 2600, 1),     # wide join part: four K-blocks
    (5, 3, 1, 300, 1),         # tiny everything: single partial tile
    (61, 151, 6, 129, 1),      # database of one tile + one row
]


@pytest.mark.parametrize("shape", SHAPES)
def test_shape_fuzz_tensor_core_vs_float64(shape):
    """Random databases of awkward shapes: the tensor-core engine (whatever path it picks) and the SIMT engine
    must both return the float64 brute-force answer in the joint and in the target space."""
    import snickery_b200
    Dt, Dj, m, N, k = shape
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 31))
    walk = np.cumsum(rng.standard_normal((N + 1, Dt + Dj)) * 0.3, axis=0)     # trajectories: neighbours cluster
    walk -= walk.mean(0)
    F = walk[:N, :Dt].astype(np.float32)
    Jc = walk[:, Dt:].astype(np.float32)
    wt = rng.uniform(0.2, 1.0, Dt)
    wj = rng.uniform(0.2, 1.0, Dj)
    db = snickery_b200.UnitDatabase(F, Jc, multiepoch=m)
    db.set_weights(wt, wj)
    Fw = F.astype(np.float64) * wt
    Jw = Jc.astype(np.float64) * wj
    Np = N - m + 1
    joint = np.hstack([Jw[:Np]] + [Fw[j:Np + j] for j in range(m)])
    nq = 70
    rows = rng.integers(0, Np, nq)
    qj = joint[rows] + 0.05 * rng.standard_normal((nq, joint.shape[1]))
    qj[:5] = joint[rows[:5]]                                                # exact hits: distance 0
    qt = Fw[rows] + 0.05 * rng.standard_normal((nq, Dt))
    for space, data, q in ((engine.SPACE_JOINT, joint, qj), (engine.SPACE_TARGET, Fw, qt)):
        kk = min(k, data.shape[0])
        rd, ri = O.brute_force_knn(data, q, kk)
        for eng in (engine.ENGINE_AUTO, engine.ENGINE_SIMT):
            db.set_engine(eng)
            d, i = db.knn(q, kk, space)
            assert_knn_matches(d, i, rd, ri)
    assert np.all(db.knn(qj[:5], 1, engine.SPACE_JOINT)[0] == 0.0)
    db.close()


def test_cluster_multicast_path(golden_epoch, monkeypatch):
    """Opt-in thread-block-cluster path (pairs of CTAs share every database tile by TMA multicast): same answers."""
    monkeypatch.setenv("SNK_TC_CLUSTER", "1")
    db = syn.make_epoch_db(n_units=20000, seed=55)
    cfg = epoch_config(tsw=(0.5, 0.5))
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    g.db.set_engine(engine.ENGINE_TC)
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(db["F"], 300, 24, seed=6)]   # 3 query tiles: odd, padded pair
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    for b in (0, 127, 128, 255, 256, 299):
        assert_greedy_path_ok(o, utts[b], paths[b], dists[b])
    assert g.db.counters()["recertified"] == 0
    q = o.combined_rep()[::97][:200] + 0.01
    rd, ri = o.joint_tree.query(q, k=12)                     # store / emit paths under the cluster launch
    d, i = g.joint_tree.query(q, k=12)
    assert_knn_matches(d, i, rd, ri)


# ------------------------------------------------------------------------------------ BASELINE.json configs[0]
def test_config1_slt_simplified_mini():
    """configs[0] (config/slt_simplified_mini.cfg:30-38,68-72,82,94): ~70k-epoch database, target weights
    [0.1, 1.0], join weights [.25]*4, join_cost_weight 0.2, multiepoch 6, three test utterances of ~650 frames,
    greedy search with search_epsilon forced to 0 -- the reference's own engine (cKDTree) on the CPU vs the GPU."""
    db = syn.make_epoch_db(n_units=70000, seed=1234 + 1)
    cfg = epoch_config(multiepoch=6, jcw=0.2, tsw=(0.1, 1.0), jsw=(0.25, 0.25, 0.25, 0.25))
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(db["F"], 3, 650, seed=101)]
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    exact = 0
    for u, p, d in zip(utts, paths, dists):
        assert len(p) == 650 // 6
        ref, rd = o.greedy_joint_search(u, return_dists=True)
        if p == ref:
            exact += 1
            np.testing.assert_allclose(d, rd, rtol=COST_RTOL)
        else:
            assert_greedy_path_ok(o, u, p, d)
    assert exact >= 2          # identical unit sequences unless a 1e-6 tie intervenes
    assert g.db.counters()["recertified"] == 0
    # the per-stream cost report the balancing loop reads (C1)
    ts, js = g.get_scores_per_stream(utts[0], paths[0])
    assert ts.shape == (108, 2) and js.shape == (107, 4) and np.all(ts >= 0) and np.all(js >= 0)


# ------------------------------------------------------------------------------------ N4: target preparation
def _unnorm_speech(F, n_utts, T, seed):
    """compose_speech-like float32 input: de-standardised frames with the unvoiced marker in the f0 column."""
    rng = np.random.default_rng(seed)
    Dt = F.shape[1]
    mean = rng.normal(size=Dt) * 3.0
    std = rng.uniform(0.5, 4.0, size=Dt)
    utts = []
    for x in syn.make_targets(F, n_utts, T, seed=seed):
        u = (x.astype(np.float64) * std + mean).astype(np.float32)
        uv = rng.random(u.shape[0]) < 0.3
        u[uv, Dt - 1] = np.float32(O.SPECIAL_UV_VALUE)     # the lf0 stream is the last column
        utts.append(u)
    return utts, mean, std


@pytest.mark.parametrize("stats_dtype", [np.float64, np.float32])
def test_prepare_targets_bit_exact(stats_dtype):
    db = syn.make_epoch_db(n_units=3000, seed=77)
    cfg = epoch_config(multiepoch=4, jcw=0.3)
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    g = Synthesiser(cfg, db["F"], db["Jc"])
    utts, mean, std = _unnorm_speech(db["F"], 3, 57, seed=5)
    mean, std = mean.astype(stats_dtype), std.astype(stats_dtype)     # the voice file holds float32 statistics
    g.set_standardisation(mean, std)
    assert O.standardise(utts[0], mean, std).dtype == stats_dtype
    for u in utts:
        ref = O.weight(O.standardise(u, mean, std), o.target_weight_vector)
        got = g.prepare_targets(u)
        assert got.dtype == np.float64 and np.array_equal(got, ref)
    assert g.prepare_targets(np.zeros((0, db["F"].shape[1]), np.float32)).shape == (0, db["F"].shape[1])


def test_greedy_from_unnormalised_speech():
    """The fused form (float32 upload, standardise + weight inside the query assembly) selects exactly what
    the reference's order of operations does: standardise -> weight -> greedy_joint_search."""
    db = syn.make_epoch_db(n_units=20000, seed=78)
    cfg = epoch_config(multiepoch=6, jcw=0.2)
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    utts, mean, std = _unnorm_speech(db["F"], 5, 90, seed=6)
    utts[1] = utts[1][:43]                                    # ragged batch
    with pytest.raises(engine.EngineError, match="standardisation"):
        g.greedy_joint_search_unnorm_batch(utts)             # standardisation not set yet
    g.set_standardisation(mean, std)
    paths, dists = g.greedy_joint_search_unnorm_batch(utts, return_dists=True)
    feats = [O.weight(O.standardise(u, mean, std), o.target_weight_vector) for u in utts]
    p64, d64 = g.greedy_joint_search_batch(feats, return_dists=True)
    assert paths == p64
    for a, b in zip(dists, d64):
        assert np.array_equal(a, b)
    for u, p, d in zip(feats, paths, dists):
        ref, rd = o.greedy_joint_search(u, return_dists=True)
        if p == ref:
            np.testing.assert_allclose(d, rd, rtol=COST_RTOL)
        else:
            assert_greedy_path_ok(o, u, p, d)
    # a re-weighting is picked up without touching the standardisation
    g.reconfigure_settings({"join_cost_weight": 0.6})
    o2 = O.OracleSynthesiser(dict(cfg, join_cost_weight=0.6), db["F"], db["Jc"])
    feats2 = [O.weight(O.standardise(u, mean, std), o2.target_weight_vector) for u in utts]
    assert g.greedy_joint_search_unnorm_batch(utts) == g.greedy_joint_search_batch(feats2)


# ------------------------------------------------------------------------------------ N3: voice file -> device
def test_synthesiser_from_voice_file(tmp_path):
    """A database dump in the reference's HDF5 layout loads into the resident database; its float32
    mean / std drive the fused float32 standardisation (what numpy does with the file's arrays)."""
    from snickery_b200.hdf5_voice import save_voice
    db = syn.make_epoch_db(n_units=8000, seed=79)
    utts, mean, std = _unnorm_speech(db["F"], 3, 60, seed=7)
    mean, std = mean.astype(np.float32), std.astype(np.float32)
    Dj = db["Jc"].shape[1]
    path = str(tmp_path / "voice.hdf5")
    save_voice(path, {"train_unit_features": db["F"], "join_contexts": db["Jc"], "mean_target": mean,
                      "std_target": std, "mean_join": np.zeros(Dj // 4, np.float32),
                      "std_join": np.ones(Dj // 4, np.float32)}, chunk_bytes=1 << 16)
    cfg = epoch_config(multiepoch=6, jcw=0.2)
    g = Synthesiser.from_voice(cfg, path)
    assert g.number_of_units == 8000 and g.mean_vec_target.dtype == np.float32
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    paths, dists = g.greedy_joint_search_unnorm_batch(utts, return_dists=True)
    for u, p, d in zip(utts, paths, dists):
        feats = O.weight(O.standardise(u, mean, std), o.target_weight_vector)     # float32 standardise, float64 weight
        assert np.array_equal(g.prepare_targets(u), feats)
        ref, rd = o.greedy_joint_search(feats, return_dists=True)
        if p == ref:
            np.testing.assert_allclose(d, rd, rtol=COST_RTOL)
        else:
            assert_greedy_path_ok(o, feats, p, d)


def test_stashable_tree_save_and_resurrect(tmp_path):
    """StashableKDTree.py:43-102: save_hdf / resurrect_tree round trip gives the same neighbours."""
    from snickery_b200.kdtree import resurrect_tree
    rng = np.random.default_rng(11)
    data = rng.normal(size=(3000, 20))
    tree = GpuStashableKDTree(data, leaf_size=100, metric="euclidean")
    q = rng.normal(size=(17, 20))
    d0, i0 = tree.query(q, k=5)
    path = str(tmp_path / "tree.hdf5")
    tree.save_hdf(path)
    again = resurrect_tree(path)
    d1, i1 = again.query(q, k=5)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    rd, ri = O.brute_force_knn(data.astype(np.float32).astype(np.float64), q, 5)
    assert_knn_matches(d0, i0, rd, ri)


def test_replicate_is2018_variant():
    """REPLICATE_IS2018_EXP (config/IS2018_nick_simplified.cfg:4; synth_simple.py:384-387): the first and the last
    target frame are dropped before the search."""
    db = syn.make_epoch_db(n_units=6000, seed=80)
    cfg = epoch_config(multiepoch=6, jcw=0.2, tsw=(0.5, 0.5))
    cfg["REPLICATE_IS2018_EXP"] = True
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    utts, mean, std = _unnorm_speech(db["F"], 2, 62, seed=8)
    g.set_standardisation(mean, std)
    paths, dists = g.greedy_joint_search_unnorm_batch(utts, return_dists=True)
    for u, p, d in zip(utts, paths, dists):
        feats = O.weight(O.standardise(u, mean, std)[1:-1, :], o.target_weight_vector)
        assert np.array_equal(g.prepare_targets(u), feats)
        assert len(p) == 60 // 6
        assert_greedy_path_ok(o, feats, p, d)
