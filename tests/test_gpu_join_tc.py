"""The tensor-core join-cost kernel (csrc/join_tc.cu) against the reference's float64 arithmetic
(get_natural_distance_vectorised, synth_halfphone.py:2942-2951) and against the direct-difference kernel it replaces
for n_candidates <= 64."""
import os

import numpy as np
import pytest

from conftest import halfphone_config
from oracle import snickery_oracle as O
from snickery_b200 import Synthesiser, synthetic as syn

pytestmark = pytest.mark.gpu

TILE_RTOL = 1e-5     # worst finite tile entry against sqrt(sum((end[a] - start[c])^2)) in float64 (rows that nearly coincide)
TILE_P99 = 6e-6      # 99 % of the entries
TIE_RTOL = 1e-6
COST_RTOL = 1e-5


@pytest.fixture(scope="module")
def voice():
    hp = syn.make_halfphone_db(n_units=20000, seed=77)
    cfg = halfphone_config(n_candidates=50, preselection="acoustic")
    o = O.OracleSynthesiser(cfg, hp["F"], hp["Jc"])
    g = Synthesiser(cfg, hp["F"], hp["Jc"])
    return hp, o, g


def reference_tiles(o, cand):
    """float64 join costs of every (a in C[t], c in C[t+1]) pair, inf where the reference enumerates no pair
    (synth_halfphone.py:3238-3268)."""
    n = o.unit_end_data.shape[0]
    T, K = cand.shape
    out = np.full((T - 1, K, K), np.inf)
    for t in range(T - 1):
        a, c = cand[t], cand[t + 1]
        oka = (a >= 1) & (a < n - 1)
        okc = (c >= 1) & (c < n - 1)
        e = o.unit_end_data[np.where(oka, a, 1)]
        s = o.unit_start_data[np.where(okc, c, 1)]
        d = np.sqrt(((e[:, None, :] - s[None, :, :]) ** 2).sum(axis=2))
        out[t] = np.where(oka[:, None] & okc[None, :], d, np.inf)
    return out


def lattices(hp, rng, T, K, n):
    """Candidate lists the way real searches produce them: neighbours of a moving target (rows near a trajectory through
    the database, so consecutive sets hold natural continuations u -> u + 1 and near-coincident rows u -> u + 2),
    plus padding, inadmissible ids and repeats."""
    start = int(rng.integers(10, n - T - 200))
    cand = np.empty((T, K), dtype=np.int64)
    for t in range(T):
        near = start + t + rng.integers(-3, 4, size=K // 2)
        far = rng.integers(0, n, size=K - K // 2)
        cand[t] = np.concatenate([near, far])
    cand[3, :4] = -1
    cand[5, 0] = 0
    cand[6, 1] = n - 1
    cand[7, K - 1] = cand[7, 0]
    return cand


@pytest.mark.parametrize("K", [50, 64, 33, 8, 3])
def test_tiles_against_float64_formula(voice, K):
    hp, o, g = voice
    rng = np.random.default_rng(K)
    n = hp["F"].shape[0]
    cands = [lattices(hp, rng, 24, K, n) for _ in range(3)]
    tiles = g.db.join_tiles(cands)
    fin_total, patched = g.db.join_stats()
    ref = np.concatenate([reference_tiles(o, c) for c in cands])
    assert tiles.shape == ref.shape
    assert np.array_equal(np.isinf(tiles), np.isinf(ref))
    fin = np.isfinite(ref)
    assert fin_total == int(fin.sum())
    zero = fin & (ref == 0)
    assert zero.any() and np.all(tiles[zero] == 0)                 # natural joins cost exactly nothing
    pos = fin & (ref > 0)
    rel = np.abs(tiles[pos] - ref[pos]) / ref[pos]
    assert rel.max() <= TILE_RTOL, "worst tile entry off by %.3g relative" % rel.max()
    assert np.quantile(rel, 0.99) <= TILE_P99
    assert 0 < patched < 0.1 * fin_total                           # near-coincident rows exist here and are few


def test_search_on_tc_tiles_equals_direct_difference_tiles(voice):
    """The same lattices searched over tensor-core tiles and over the fp32 direct-difference tiles they replaced."""
    hp, o, g = voice
    rng = np.random.default_rng(5)
    n = hp["F"].shape[0]
    cands, dists = [], []
    for b in range(6):
        T = int(rng.integers(2, 90))
        c = lattices(hp, rng, max(T, 8), 50, n)[:T]
        cands.append(c)
        dists.append(rng.random((T, 50)) * 3.0)
    fused = g.viterbi_search_batch(cands, dists, return_costs=True)
    os.environ["SNK_JOIN_NOTC"] = "1"
    try:
        plain = g.viterbi_search_batch(cands, dists, return_costs=True)
    finally:
        del os.environ["SNK_JOIN_NOTC"]
    for b in range(6):
        ref_path, ref_cost = O.viterbi_search_numpy(o, cands[b], dists[b], return_cost=True)
        for paths, pc, tc, jc in (fused, plain):
            if not ref_path:
                assert paths[b] == []
                continue
            _, _, tot = o.path_costs(cands[b], dists[b], paths[b])
            assert tot <= ref_cost * (1 + TIE_RTOL)
            if paths[b] != ref_path:
                assert abs(tot - ref_cost) <= TIE_RTOL * ref_cost
            assert abs(pc[b] - tot) <= COST_RTOL * tot
        assert abs(fused[3][b] - plain[3][b]) <= COST_RTOL * max(plain[3][b], 1e-12)      # join cost of the path


def test_reweighting_moves_operand_scale(voice):
    hp, o, g = voice
    rng = np.random.default_rng(9)
    cand = lattices(hp, rng, 12, 50, hp["F"].shape[0])
    before = g.db.join_tiles([cand])
    w = g.join_weight_vector.copy()
    g.db.set_weights(g.target_weight_vector, w * 37.5)           # also moves the power-of-two operand scale
    try:
        after = g.db.join_tiles([cand])
    finally:
        g.db.set_weights(g.target_weight_vector, w)
    fin = np.isfinite(before)
    np.testing.assert_allclose(after[fin], before[fin] * 37.5, rtol=TILE_RTOL)
    np.testing.assert_array_equal(g.db.join_tiles([cand]), before)
