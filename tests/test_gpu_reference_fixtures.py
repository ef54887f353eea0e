"""GPU parity against outputs of the reference's OWN code (tests/golden/reference_exec.npz, produced by
tests/golden/make_reference_fixtures.py through oracle/ref_exec.py), and at-size parity against the reference's own
k-NN engine (scipy cKDTree) / sklearn KDTree.  Everything goes through the C ABI.

Tolerances (BASELINE.json): unit sequences bit-exact except at cost ties within 1e-6 relative; costs within 1e-5."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_reference_fixtures as MF  # noqa: E402
from conftest import GOLDEN, epoch_config, halfphone_config  # noqa: E402
from oracle import snickery_oracle as O  # noqa: E402
from snickery_b200 import GpuStashableKDTree, Synthesiser, engine, synthetic as syn  # noqa: E402
from test_gpu_parity import COST_RTOL, TIE_RTOL, assert_greedy_path_ok, assert_knn_matches, check_viterbi  # noqa: E402

pytestmark = pytest.mark.gpu
ENGINES = [engine.ENGINE_SIMT, engine.ENGINE_TC]


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "reference_exec.npz"))


@pytest.fixture(scope="module")
def inputs(fx):
    return {k[3:]: fx[k] for k in fx.files if k.startswith("in_")}


# ------------------------------------------------------------------------------------ reference-executed fixtures
@pytest.mark.parametrize("eng", ENGINES)
def test_greedy_equals_reference_run(fx, golden_epoch, eng):
    """synth_simple.Synthesiser.greedy_joint_search as the reference computes it (its own cKDTree, eps = 0)."""
    ge = golden_epoch
    for tag, cfg in MF.epoch_cases().items():
        g = Synthesiser(cfg, ge["F"], ge["Jc"])
        g.db.set_engine(eng)
        assert np.array_equal(g.target_weight_vector, fx["%s_wt" % tag])
        for i in MF.epoch_targets(tag):
            uf = MF.epoch_unit_features(cfg, ge, i, g.target_weight_vector, getattr(g, "target_truncation_vector", None))
            assert g.greedy_joint_search(uf) == fx["%s_path_%d" % (tag, i)].tolist(), (tag, i)
    g = Synthesiser(MF.epoch_cases()["cfg1_m6"], ge["F"], ge["Jc"])
    g.db.set_engine(eng)
    tf = ge["F"][400:400 + 72].astype(np.float64) * g.target_weight_vector
    assert g.greedy_joint_search(tf, start_state=400) == fx["identity_path"].tolist()


@pytest.mark.parametrize("eng", ENGINES)
def test_halfphone_epoch_layout_and_scores_equal_reference_run(fx, golden_epoch, eng):
    ge = golden_epoch
    for m in (1, 3):
        cfg = dict(epoch_config(multiepoch=m), halfphone_epoch_join_layout=True)
        g = Synthesiser(cfg, ge["F"], MF.hp_epoch_join(ge["Jc"]))
        g.db.set_engine(eng)
        uf = MF.epoch_unit_features(cfg, ge, 1, g.target_weight_vector, None)[:60]
        p = g.greedy_joint_search(uf)
        assert p == fx["hpepoch_m%d_path" % m].tolist()
        ts, js = g.get_scores_per_stream(uf, p)
        np.testing.assert_allclose(js, fx["hpepoch_m%d_jscores" % m], rtol=1e-12, atol=1e-300)
        if m == 1:      # for m > 1 the reference slices only the first frame's columns; scores.cu sums the window
            np.testing.assert_allclose(ts, fx["hpepoch_m%d_tscores" % m], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("eng", ENGINES)
def test_acoustic_preselection_and_viterbi_equal_reference_run(fx, golden_halfphone, eng):
    """preselect_units_acoustic (cKDTree k-NN) -> viterbi_search (T o J, shortest path) as the reference runs them."""
    gh = golden_halfphone
    for K in (12, 50):
        g = Synthesiser(halfphone_config(n_candidates=K), gh["F"], gh["Jc"])
        g.db.set_engine(eng)
        cand, dist = g.preselect_units_acoustic(gh["targets"])
        assert_knn_matches(dist, cand, fx["hp_k%d_dist" % K], fx["hp_k%d_cand" % K])
        assert np.array_equal(cand, fx["hp_k%d_cand" % K])
        paths, pc, tc, jc = g.viterbi_search_batch([fx["hp_k%d_cand" % K]], [fx["hp_k%d_dist" % K]], return_costs=True)
        assert paths[0] == fx["hp_k%d_path" % K].tolist()
        ref_cost = float(fx["hp_k%d_cost" % K])
        assert abs(pc[0] - ref_cost) <= COST_RTOL * ref_cost
        assert g.viterbi_search(cand, dist) == paths[0]


def test_lattice_semantics_equal_reference_run(fx, inputs, golden_halfphone):
    """-1 padding, unit 0 / N-1 never joinable, duplicates, blocked frames, natural joins, K in {1, 30, 64}."""
    gh = golden_halfphone
    g = Synthesiser(halfphone_config(n_candidates=12), gh["F"], gh["Jc"])
    names = MF.lattice_names()
    for name in names:
        cand, dist = inputs[name + "_cand"], inputs[name + "_dist"]
        paths, pc, _, _ = g.viterbi_search_batch([cand], [dist], return_costs=True)
        assert paths[0] == fx[name + "_path"].tolist(), name
        if paths[0]:
            ref_cost = float(fx[name + "_cost"])
            assert abs(pc[0] - ref_cost) <= COST_RTOL * ref_cost
    tiles = g.db.join_tiles([inputs["lat_natural_cand"][:2]])
    np.testing.assert_allclose(tiles[0], fx["lat_natural_tile0"], rtol=COST_RTOL, atol=0)
    assert tiles[0][0, 0] == 0.0 and fx["lat_natural_tile0"][0, 0] == 0.0


def test_label_preselection_equals_reference_run(fx, golden_halfphone):
    gh = golden_halfphone
    names = MF.halfphone_names(gh["phones"].tolist())
    tnames = [names[i] for i in (100, 101, 102, 300, 301, 302, 640, 641)]
    g = Synthesiser(halfphone_config(n_candidates=12, preselection="quinphone"), gh["F"], gh["Jc"], train_unit_names=names)
    cq, dq = g.preselect_units_quinphone(gh["targets"][:8], tnames)
    assert np.array_equal(cq, fx["quin_cand"])
    np.testing.assert_allclose(dq, fx["quin_dist"], rtol=1e-12)
    assert g.viterbi_search(cq, dq) == fx["quin_path"].tolist()
    gm = Synthesiser(halfphone_config(n_candidates=6, preselection="monophone_then_acoustic"), gh["F"], gh["Jc"],
                     train_unit_names=names)
    cm, dm = gm.preselect_units_monophone_then_acoustic(gh["targets"][:8], tnames)
    assert np.array_equal(cm, fx["mono_cand"])
    np.testing.assert_allclose(dm, fx["mono_dist"], rtol=1e-9)


def test_target_preparation_equals_reference_run(fx, inputs, golden_epoch):
    """data_manipulation.standardise + speech_manip.weight as the reference's functions return them, bit for bit."""
    ge = golden_epoch
    cfg = epoch_config()
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        g = Synthesiser(cfg, ge["F"], ge["Jc"])
        w = np.linspace(0.1, 1.0, 61)
        g.db.set_weights(w, g.join_weight_vector)
        g.db.set_standardisation(inputs["std_mean"].astype(cast), inputs["std_std"].astype(cast).reshape(-1))
        out = g.db.prepare_targets(inputs["std_speech"])
        assert np.array_equal(out, np.asarray(fx["std_weighted_" + nm], dtype=np.float64))


# ------------------------------------------------------------------------------------ at size, against the reference's engines
def test_config2_full_size_vs_ckdtree():
    """BASELINE.json configs[1] at full size: 700k-unit database, one 648-frame utterance = 108 greedy steps,
    against the reference's own engine -- scipy cKDTree(leafsize=100, balanced_tree=False) over the 517-dim joint
    rows, eps = 0 (synth_simple.py:229,490) -- driven by the oracle's restatement of the chain."""
    import bench
    cfg = bench.workload_config()
    db = bench.make_database(bench.DB_UNITS)
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    g.db.set_engine(engine.ENGINE_TC)
    cat = bench.make_batch(db["F"], g.target_weight_vector, 2, bench.UTT_FRAMES, seed=777)
    utts = [cat[:bench.UTT_FRAMES], cat[bench.UTT_FRAMES:bench.UTT_FRAMES + 6 * 20]]      # 108 + 20 steps
    paths, dists = g.greedy_joint_search_batch(utts, return_dists=True)
    assert len(paths[0]) == 108
    mism = 0
    for u, p, d in zip(utts, paths, dists):
        ref, rd = o.greedy_joint_search(u, return_dists=True)
        if p == ref:
            np.testing.assert_allclose(d, rd, rtol=COST_RTOL, atol=1e-12)
        else:           # a tie within 1e-6 may send the two chains apart: audit step by step
            mism += assert_greedy_path_ok(o, u, p, d)
    assert mism <= 2
    assert g.db.counters()["recertified"] == 0
    # the reference's literal call, one utterance: the persistent single-utterance kernel (greedy_one.cu) at full size --
    # the same path and bit-identical distances as the batched tensor-core path just checked against cKDTree
    g.db.counters(reset=True)
    p1, d1 = g.greedy_joint_search_batch(utts[:1], return_dists=True)
    assert g.db.counters()["launches"] <= 3 and g.db.counters()["recertified"] == 0
    assert p1[0] == paths[0] and np.array_equal(np.asarray(d1[0]), np.asarray(dists[0]))


def test_config3_full_size_vs_reference_engines():
    """BASELINE.json configs[2] at size: 90k half-phones, 184-dim targets, K = 50, T = 80: acoustic preselection against
    the reference's cKDTree(leafsize=100, compact_nodes=False, balanced_tree=False).query(k=50)
    (synth_halfphone.py:379,1364), then join costs + Viterbi against the oracle's DP (pinned by the reference run)."""
    hp = syn.make_halfphone_db(n_units=90000, seed=1237)
    cfg = halfphone_config(n_candidates=50, preselection="acoustic")
    o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
    o.build_acoustic_tree()
    g = Synthesiser(cfg, hp["F"], hp["Jc"])
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(hp["F"], 3, 80, seed=31)]
    cands, dists = [], []
    for uf in utts:
        c, d = g.preselect_units_acoustic(uf)
        rc, rd = o.preselect_units_acoustic(uf)
        assert_knn_matches(d, c, rd, rc)
        cands.append(c)
        dists.append(d)
    assert g.db.counters()["recertified"] == 0
    paths, pc, tc, jc = g.viterbi_search_batch(cands, dists, return_costs=True)
    for b in range(3):
        assert len(paths[b]) == 80
        check_viterbi(o, cands[b], dists[b], paths[b], pc[b], tc[b], jc[b])
    # quinphone-style lattices at the same size: label lookups (host) + GPU distances + Viterbi
    cq = syn.quinphone_like_candidates(hp["phones"], hp["phones"][5000:5080], 50, seed=3)
    dq = g.candidate_target_distances(cq, utts[0])
    np.testing.assert_allclose(dq, o.candidate_distances(cq, utts[0]), rtol=1e-12)
    p, pcq, tcq, jcq = g.viterbi_search_batch([cq], [dq], return_costs=True)
    check_viterbi(o, cq, dq, p[0], pcq[0], tcq[0], jcq[0])


def test_stashable_tree_vs_sklearn_kdtree():
    """S1: StashableKDTree is sklearn.neighbors.KDTree(data, leaf_size=100, metric='euclidean') (StashableKDTree.py:7,
    active_learning_join.py:198-202): `.query(X, k)` must return what sklearn returns."""
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(8)
    data = rng.standard_normal((20000, 20)).astype(np.float32).astype(np.float64)     # join vectors come from float32 files
    X = data[rng.integers(0, 20000, size=300)] + 0.05 * rng.standard_normal((300, 20))
    ref = KDTree(data, leaf_size=100, metric="euclidean")
    tree = GpuStashableKDTree(data, leaf_size=100, metric="euclidean")
    for k in (1, 5, 40):
        rd, ri = ref.query(X, k=k)
        d, i = tree.query(X, k=k)
        assert d.shape == rd.shape == (300, k) and i.dtype == np.int64
        assert_knn_matches(d, i, rd, ri)
    assert np.array_equal(tree.query(data[:50], k=1)[1][:, 0], np.arange(50))        # a stored point finds itself
    ii = tree.query(X, k=3, return_distance=False)
    assert np.array_equal(ii, ref.query(X, k=3, return_distance=False))


def test_halfphone_target_preparation_equals_reference_run(fx, inputs, golden_halfphone):
    """Row N4, half-phone part: standardise -> get_halfphone_stats (three-point sampling from the state alignment) ->
    hstack(durations) -> weight (synth_halfphone.py:1510-1548, train_halfphone.py:959-1070), bit for bit in both of
    numpy's arithmetic modes, from un-normalised float32 speech in one device kernel."""
    gh = golden_halfphone
    labs = MF.hp3_labels(inputs["hp3_state_ends"])
    g = Synthesiser(halfphone_config(n_candidates=4), gh["F"], gh["Jc"])
    assert g.db.Dt == 184 and g.target_representation == "threepoint"
    g.db.set_weights(np.linspace(0.2, 1.1, 184), g.join_weight_vector)
    g._dirty = False
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        g.set_standardisation(inputs["std_mean"].astype(cast), inputs["std_std"].astype(cast).reshape(-1))
        names, feats, timings = g.halfphone_targets(inputs["hp3_speech"], labs, durations=inputs["hp3_dur"])
        assert names.tolist() == fx["hp3_names"].tolist()
        assert np.array_equal(np.asarray(timings), fx["hp3_timings"])
        assert np.array_equal(feats, fx["hp3_threepoint_dur_" + nm])
    with pytest.raises(ValueError):
        g.halfphone_targets(inputs["hp3_speech"], labs)                 # 183 columns, the voice has 184
