"""CPU tests of the oracle: pins it to the reference's own engines / known answers and to the
committed golden vectors."""
import numpy as np
import pytest

from conftest import epoch_config, halfphone_config
from oracle import snickery_oracle as O
from snickery_b200 import synthetic as syn


@pytest.fixture(scope="module")
def epoch_synth(golden_epoch):
    s = O.OracleSynthesiser(epoch_config(), golden_epoch["F"], golden_epoch["Jc"])
    s.get_tree_for_greedy_search()
    return s


@pytest.fixture(scope="module")
def hp_synth(golden_halfphone):
    s = O.OracleSynthesiser(halfphone_config(n_candidates=12), golden_halfphone["F"], golden_halfphone["Jc"])
    s.build_acoustic_tree()
    return s


def test_weighting_is_f32_times_f64(epoch_synth, golden_epoch):
    # speech_manip.py:209-213: float32 data * float64 weights -> float64
    assert epoch_synth.train_unit_features.dtype == np.float64
    w = epoch_synth.target_weight_vector
    assert w.shape == (61,) and np.allclose(w[:60], 0.1 * 0.8) and np.isclose(w[60], 1.0 * 0.8)
    assert np.array_equal(epoch_synth.train_unit_features,
                          golden_epoch["F"].astype(np.float64) * w[None, :])
    assert np.allclose(epoch_synth.join_weight_vector, 0.25 * 0.2)


def test_natural_neighbours_join_at_zero(epoch_synth):
    # synth_simple.py:250-251: unit_end_data[u] is unit_start_data[u+1]
    assert np.array_equal(epoch_synth.unit_end_data[:-1], epoch_synth.unit_start_data[1:])
    d = epoch_synth.get_natural_distance_vectorised(np.arange(10, 50), np.arange(11, 51), order=1)
    assert np.all(d == 0.0)


def test_multiepoch_layout(epoch_synth):
    m, n, dt = 6, 1500, 61
    c = epoch_synth.combined_rep()
    assert c.shape == (n - m + 1, 151 + m * dt)
    u = 77
    assert np.array_equal(c[u, :151], epoch_synth.unit_start_data[u])
    for j in range(m):
        assert np.array_equal(c[u, 151 + j * dt:151 + (j + 1) * dt], epoch_synth.train_unit_features[u + j])
    assert np.array_equal(epoch_synth.current_join_rep[u], epoch_synth.unit_end_data[u + m - 1])


def test_identity_path_known_answer(epoch_synth, golden_epoch):
    # the reference's only known answer (synth_simple.py:909-928), m > 1 form
    start, m, n = int(golden_epoch["identity_start"]), 6, 12
    tf = epoch_synth.train_unit_features[start:start + m * n]
    for engine in ("tree", "brute"):
        assert epoch_synth.greedy_joint_search(tf, start_state=start, engine=engine) == \
            list(range(start, start + m * n, m))
    assert golden_epoch["identity_path"].tolist() == list(range(start, start + m * n, m))


def test_greedy_tree_equals_bruteforce_and_golden(epoch_synth, golden_epoch):
    for i in range(3):
        uf = golden_epoch["targets_%d" % i]
        p_tree, d_tree = epoch_synth.greedy_joint_search(uf, return_dists=True)
        p_brute, d_brute = epoch_synth.greedy_joint_search(uf, engine="brute", return_dists=True)
        assert p_tree == p_brute == golden_epoch["path_%d" % i].tolist()
        assert len(p_tree) == 121 // 6
        np.testing.assert_allclose(d_tree, golden_epoch["dist_%d" % i], rtol=1e-12)
        np.testing.assert_allclose(d_tree, d_brute, rtol=1e-12)


def test_greedy_short_utterance_raises(epoch_synth, golden_epoch):
    with pytest.raises(ValueError):
        epoch_synth.greedy_joint_search(golden_epoch["targets_0"][:5])   # fewer than multiepoch frames


def test_acoustic_knn_tree_equals_bruteforce_and_golden(hp_synth, golden_halfphone):
    uf = golden_halfphone["targets"]
    cand, dist = hp_synth.preselect_units_acoustic(uf)
    dd, ii = O.brute_force_knn(hp_synth.train_unit_features, uf, 12)
    assert np.array_equal(cand, ii) and np.array_equal(cand, golden_halfphone["knn_idx"])
    np.testing.assert_allclose(dist, dd, rtol=1e-12)
    np.testing.assert_allclose(dist, golden_halfphone["knn_dist"], rtol=1e-12)
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_viterbi_dp_equals_exhaustive_enumeration(hp_synth, golden_halfphone):
    rng = np.random.default_rng(0)
    n = hp_synth.unit_end_data.shape[0]
    for trial in range(6):
        T, K = int(rng.integers(2, 6)), int(rng.integers(2, 5))
        cand = rng.integers(-1, n, size=(T, K))
        if trial % 2:
            cand[rng.integers(0, T), rng.integers(0, K)] = 0        # inadmissible first unit
            cand[rng.integers(0, T), rng.integers(0, K)] = n - 1    # inadmissible last unit
        dist = rng.random((T, K))
        pe, ce = hp_synth.viterbi_exhaustive(cand, dist)
        pv, cv = hp_synth.viterbi_search(cand, dist, return_cost=True)
        pn, cn = O.viterbi_search_numpy(hp_synth, cand, dist, return_cost=True)
        if not pe:
            assert pv == [] and pn == []
            continue
        assert abs(ce - cv) < 1e-9 and abs(ce - cn) < 1e-9
        assert pe == pv == pn


def test_viterbi_golden_and_variants(hp_synth, golden_halfphone):
    g = golden_halfphone
    p, c = hp_synth.viterbi_search(g["knn_idx"], g["knn_dist"], return_cost=True)
    assert p == g["vit_path_f64"].tolist() and abs(c - float(g["vit_cost_f64"])) < 1e-9
    p32, c32 = hp_synth.viterbi_search(g["knn_idx"], g["knn_dist"], arithmetic="openfst32", return_cost=True)
    assert p32 == g["vit_path_fst32"].tolist()
    assert abs(c32 - c) <= 1e-5 * c     # float32 accumulation stays within the cost tolerance
    pq, cq = hp_synth.viterbi_search(g["q_cand"], g["q_dist"], return_cost=True)
    assert pq == g["q_path"].tolist() and abs(cq - float(g["q_cost"])) < 1e-9
    tc, jc, tot = hp_synth.path_costs(g["q_cand"], g["q_dist"], pq)
    assert abs(tot - cq) < 1e-9


def test_viterbi_quirks(hp_synth):
    n = hp_synth.unit_end_data.shape[0]
    d = np.ones((3, 2))
    # single frame -> empty J -> empty path (fst_functions_wrapped.py:172-217)
    assert hp_synth.viterbi_search(np.array([[5, 6]]), np.ones((1, 2))) == []
    # unit 0 and unit N-1 can never be used (synth_halfphone.py:3238-3240)
    assert hp_synth.viterbi_search(np.array([[0, 0], [5, 6], [7, 8]]), d) == []
    assert hp_synth.viterbi_search(np.array([[4, 5], [n - 1, n - 1], [7, 8]]), d) == []
    # -1 padding is skipped, duplicates are harmless
    p = hp_synth.viterbi_search(np.array([[4, -1], [5, 5], [-1, 6]]), d)
    assert p == [4, 5, 6]
    # natural continuation costs nothing to join
    p, c = hp_synth.viterbi_search(np.array([[4, 900], [5, 901], [6, 902]]), np.zeros((3, 2)), return_cost=True)
    assert c == 0.0 and p in ([4, 5, 6], [900, 901, 902])


def test_candidate_distances_negative_index(hp_synth, golden_halfphone):
    g = golden_halfphone
    d = hp_synth.candidate_distances(g["q_cand"], g["targets"])
    np.testing.assert_allclose(d, g["q_dist"], rtol=1e-12)
    t, j = np.argwhere(g["q_cand"] == -1)[0]
    last = hp_synth.train_unit_features[-1]
    assert np.isclose(d[t, j], np.sqrt(((last - g["targets"][t]) ** 2).sum()))


def test_per_stream_scores_golden(epoch_synth, golden_epoch):
    uf = golden_epoch["targets_0"]
    p = golden_epoch["path_0"].tolist()
    ts = epoch_synth.get_target_scores_per_stream(epoch_synth.window_targets(uf), p)
    js = epoch_synth.get_join_scores_per_stream(p)
    # aggregate_squared_errors_by_stream only walks the first Dt columns of a windowed row
    np.testing.assert_allclose(ts, golden_epoch["tscores_0"], rtol=1e-12)
    np.testing.assert_allclose(js, golden_epoch["jscores_0"], rtol=1e-12)
    assert js.shape == (len(p) - 1, 4) and ts.shape == (len(p), 2)


def test_synthetic_generator_shapes_and_ties():
    db = syn.make_epoch_db(n_units=900, seed=3)
    assert db["F"].shape == (900, 61) and db["Jc"].shape == (901, 151)
    assert np.array_equal(db["Jc"][0], db["Jc"][1])               # duplicated first frame (train_simple.py:215-217)
    assert np.array_equal(db["F"][:, :60], db["Jc"][1:, :60])     # same mag stream feeds both costs
    uv = db["F"][:, 60] == syn.UV_VALUE
    assert 0.05 < uv.mean() < 0.7                                  # exact-tie unvoiced runs exist
    hp = syn.make_halfphone_db(n_units=500, seed=4)
    assert hp["F"].shape == (500, 184) and hp["Jc"].shape == (501, 151)


def test_magphase_concatenation_oracle_known_answers():
    """Row N2 restatement (synth_simple.py:538-747): a natural run of units cross-fades back to the original
    frames (in-taper + out-taper = 1), edges are zero padded, unvoiced f0 is zeroed."""
    rng = np.random.default_rng(1)
    n, w, m = 40, 17, 3
    f0 = np.where(np.arange(n)[:, None] % 10 < 6, 120.0, 0.0).astype(np.float32)
    sent = {"a": tuple(rng.standard_normal((n, w)).astype(np.float32) for _ in range(3)) + (f0,)}
    o = O.OracleMagPhaseStore(sent, ["a"] * n, np.arange(n), multiepoch=m)
    path = list(range(6, 30, m))                                # natural continuation
    for overlap in (0, 2):
        mag, real, imag, fz = o.concatenate(path, overlap=overlap)
        assert mag.shape == (len(path) * m, w) and fz.shape == (len(path) * m, 1)
        e = overlap // 2      # the very first / last kept frames only see one side of the cross-fade
        sl = slice(e, len(path) * m - e)
        np.testing.assert_allclose(mag[sl], sent["a"][0][6:30][sl], rtol=2e-7, atol=1e-7)
        np.testing.assert_allclose(imag[sl], sent["a"][2][6:30][sl], rtol=2e-7, atol=1e-7)
        voiced = f0[6:30, 0] > 0
        assert np.all(fz[sl][~voiced[sl]] == 0.0) and np.allclose(fz[sl][voiced[sl]], 120.0)
    t = np.hanning(((2 + 1) * 2) + 1)[1:3]
    assert np.allclose(t + t[::-1], 1.0)                         # matrix_operations.py:25 "check sum to 1"
    # first unit of a sentence with overlap: the leading extra frame is zero padding (then trimmed)
    mag, _, _, _ = o.concatenate([0, 20], overlap=2)
    frag = o.retrieve_magphase_frag(0, extra_frames=1)[0]
    assert frag.shape == (m + 2, w) and np.all(frag[0] == 0.0) and frag.dtype == np.float64
    with pytest.raises(AssertionError):
        o.concatenate(path, overlap=1)


def test_standardise_known_answer():
    """data_manipulation.py:162-186 by hand: plain columns are (x - mean) / std, the unvoiced marker becomes
    -20 * std whatever the mean, and the input is left untouched."""
    x = np.array([[1.0, 5.0], [3.0, O.SPECIAL_UV_VALUE], [5.0, 7.0]], dtype=np.float32)
    keep = x.copy()
    y = O.standardise(x, np.array([3.0, 6.0]), np.array([[2.0, 0.5]]))
    assert y.dtype == np.float64
    np.testing.assert_array_equal(y, np.array([[-1.0, -2.0], [0.0, -10.0], [1.0, 2.0]]))
    np.testing.assert_array_equal(x, keep)
    w = O.weight(y, [2.0, 0.5])
    np.testing.assert_array_equal(w, np.array([[-2.0, -1.0], [0.0, -5.0], [2.0, 1.0]]))
