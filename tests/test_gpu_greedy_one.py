"""The single-utterance greedy kernel (snickery_b200/csrc/greedy_one.cu): greedy_joint_search(one utterance)
(script/synth_simple.py:413, 458-503) as one persistent kernel.  It must return what the batched path returns -- bit for bit,
both re-rank in the reference's float64 arithmetic -- and what the oracle returns; its mma.sync keys must stay inside the
slack the certificate allows; an uncertified step must be repaired by the engine chain."""
import numpy as np
import pytest

from conftest import epoch_config
from oracle import snickery_oracle as O
from snickery_b200 import Synthesiser, engine, synthetic as syn
from test_gpu_parity import COST_RTOL, assert_greedy_path_ok

pytestmark = pytest.mark.gpu


def _pair(m, n_units=30000, seed=77, **kw):
    db = syn.make_epoch_db(n_units=n_units, seed=seed)
    cfg = dict(epoch_config(multiepoch=m, tsw=(0.5, 0.5)), **kw)
    o = O.OracleSynthesiser(cfg, db["F"], db["Jc"])
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, db["F"], db["Jc"])
    return db, o, g


def _launches(g, fn):
    g.db.counters(reset=True)
    out = fn()
    return out, g.db.counters()["launches"]


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 7])
def test_one_kernel_equals_batched_path_and_oracle(m, monkeypatch):
    db, o, g = _pair(m)
    steps = 24
    utts = [O.weight(x, o.target_weight_vector) for x in syn.make_targets(db["F"], 3, steps * m, seed=11 + m)]
    utts.append(o.train_unit_features[2000:2000 + steps * m])          # natural run: zero-distance answers
    for i, uf in enumerate(utts):
        start = 2000 if i == 3 else -1                                 # prev_join_rep[2000] is row 2000's own join context
        (p1, d1), n1 = _launches(g, lambda: g.greedy_joint_search_batch([uf], [start], return_dists=True))
        monkeypatch.setenv("SNK_GREEDY_NO_ONE", "1")
        (p2, d2), n2 = _launches(g, lambda: g.greedy_joint_search_batch([uf], [start], return_dists=True))
        monkeypatch.delenv("SNK_GREEDY_NO_ONE")
        assert n1 <= 3 and n2 >= 2 * steps, (n1, n2)                   # one persistent kernel against launches per step
        assert len(p1[0]) == steps
        assert p1[0] == p2[0]
        assert np.array_equal(np.asarray(d1[0]), np.asarray(d2[0]))    # the same float64 re-rank arithmetic
        assert_greedy_path_ok(o, uf, p1[0], d1[0], start_state=start)
        if i == 3:
            assert p1[0] == list(range(2000, 2000 + steps * m, m)) and np.all(np.asarray(d1[0]) == 0.0)
    assert g.db.counters()["recertified"] == 0
    # the literal call site
    assert g.greedy_joint_search(utts[0]) == g.greedy_joint_search_batch([utts[0]])[0]


def test_one_kernel_small_and_odd_sizes():
    """Databases smaller than one 16-row group per CTA, a single step, a last group that is cut."""
    for n_units in (70, 1000, 2377):
        db, o, g = _pair(6, n_units=n_units, seed=5)
        for T in (6, 13, 60):
            uf = O.weight(syn.make_targets(db["F"], 1, T, seed=T)[0], o.target_weight_vector)
            (p, d), n = _launches(g, lambda: g.greedy_joint_search_batch([uf], return_dists=True))
            assert n <= 3
            ref, rd = o.greedy_joint_search(uf, return_dists=True)
            assert len(p[0]) == T // 6
            if p[0] == ref:
                np.testing.assert_allclose(d[0], rd, rtol=COST_RTOL, atol=1e-12)
            else:
                assert_greedy_path_ok(o, uf, p[0], d[0])


def test_one_kernel_halfphone_epoch_join_layout():
    F = syn.make_epoch_db(n_units=8000, seed=3)
    Jc1 = F["Jc"]
    Jc = np.ascontiguousarray(np.hstack([Jc1, np.vstack([Jc1[1:], Jc1[-1:]])]))
    cfg = dict(epoch_config(multiepoch=3), halfphone_epoch_join_layout=True)
    o = O.OracleSynthesiser(cfg, F["F"], Jc)
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, F["F"], Jc)
    uf = O.weight(syn.make_targets(F["F"], 1, 45, seed=8)[0], o.target_weight_vector)
    (p, d), n = _launches(g, lambda: g.greedy_joint_search_batch([uf], return_dists=True))
    assert n <= 3 and len(p[0]) == 15
    assert_greedy_path_ok(o, uf, p[0], d[0])


def test_one_kernel_from_unnormalised_speech():
    from test_gpu_parity import _unnorm_speech
    db, o, g = _pair(6, n_units=20000, seed=9)
    utts, mean, std = _unnorm_speech(db["F"], 1, 60, seed=6)
    g.set_standardisation(mean, std)
    (p, d), n = _launches(g, lambda: g.greedy_joint_search_unnorm_batch(utts, return_dists=True))
    assert n <= 3
    uf = O.weight(O.standardise(utts[0], mean, std), o.target_weight_vector)
    p2, d2 = g.greedy_joint_search_batch([uf], return_dists=True)
    assert p[0] == p2[0] and np.array_equal(np.asarray(d[0]), np.asarray(d2[0]))
    assert_greedy_path_ok(o, uf, p[0], d[0])


def test_one_kernel_certificate_failure_is_repaired(monkeypatch):
    monkeypatch.setenv("SNK_DEBUG_CERT_FAIL", "1")
    db, o, g = _pair(6, n_units=12000, seed=21)
    uf = O.weight(syn.make_targets(db["F"], 1, 36, seed=4)[0], o.target_weight_vector)
    p, d = g.greedy_joint_search_batch([uf], return_dists=True)
    assert g.db.counters()["recertified"] >= 1
    assert_greedy_path_ok(o, uf, p[0], d[0])


def test_one_kernel_duplicates_and_ties_go_through_the_chain():
    """Hundreds of exact duplicates of the best row: more ties than the kernel re-ranks, so the certificate must fail and the
    chain (fp32 engine, then the exhaustive float64 scan) must return the lowest row id, as a float64 brute force does."""
    db = syn.make_epoch_db(n_units=6000, seed=13)
    F, Jc = db["F"].copy(), db["Jc"].copy()
    F[1000:1400] = F[1000]
    Jc[1000:1401] = Jc[1000]
    cfg = epoch_config(multiepoch=1, tsw=(0.5, 0.5))
    o = O.OracleSynthesiser(cfg, F, Jc)
    o.get_tree_for_greedy_search()
    g = Synthesiser(cfg, F, Jc)
    uf = O.weight(F[1000:1003], o.target_weight_vector)
    p, d = g.greedy_joint_search_batch([uf], [1000], return_dists=True)
    assert p[0][0] == 1000 and d[0][0] == 0.0
    assert g.db.counters()["recertified"] >= 1
    assert_greedy_path_ok(o, uf, p[0], d[0], start_state=1000)


@pytest.mark.parametrize("m", [6, 1])
def test_one_kernel_measured_key_error_is_inside_the_certified_slack(m):
    """The certificate's slack was derived for 34 tcgen05 K = 16 steps; the kernel runs the same 34 steps as mma.sync.m16n8k16.
    Measure its keys against float64 arithmetic on the same fp16-rounded operands."""
    db = syn.make_epoch_db(n_units=12000, seed=77)
    g = Synthesiser(epoch_config(multiepoch=m, tsw=(0.5, 0.5)), db["F"], db["Jc"])
    Fw = db["F"].astype(np.float64) * g.target_weight_vector
    Jw = db["Jc"].astype(np.float64) * g.join_weight_vector
    n = db["F"].shape[0] - (m - 1)
    rows = np.hstack([Jw[:n]] + [Fw[j:n + j] for j in range(m)])
    r16 = lambda x: np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)
    yr = r16(rows)
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(6):
        start = int(rng.integers(0, n - m - 1))
        u = int(rng.integers(0, n - m))
        window = Fw[u:u + m] + (0.05 * rng.standard_normal((m, Fw.shape[1])) * np.abs(Fw).mean() if trial % 2 else 0.0)
        keys, qn, eps_rel, maxn = g.db.debug_greedy_one_keys(window, start)
        # the query the kernel builds: prev_join_rep[start] || window
        q = np.concatenate([Jw[start], window.reshape(-1)])
        xr = r16(q)
        d2 = (xr * xr).sum() + (yr * yr).sum(1) - 2.0 * yr @ xr
        err = np.abs(keys.astype(np.float64) + float(qn) - d2)
        worst = max(worst, float(err.max() / (float(qn) + 2.0 * float(maxn))))
        assert abs(float(qn) - (xr * xr).sum()) <= 1e-5 * (xr * xr).sum()
    assert worst <= eps_rel, (worst, eps_rel)
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "cert_measured.jsonl"), "a") as f:
            f.write(json.dumps({"space": "greedy_one_m%d" % m, "operand_columns": int(rows.shape[1]), "pairs": int(6 * n),
                                "max_key_error_over_scale": worst, "certified_eps_rel": float(eps_rel)}) + "\n")
