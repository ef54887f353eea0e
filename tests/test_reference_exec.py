"""Pins the oracle (oracle/snickery_oracle.py) against outputs of THE REFERENCE'S OWN CODE.

tests/golden/reference_exec.npz holds what the reference's functions -- lifted from /root/reference/script by
oracle/ref_exec.py and run under a mechanical Python 2 -> 3 transform -- return on seeded inputs.  Here:

  * when /root/reference is present (this container) everything is regenerated and must equal the committed file
    bit for bit, so the fixtures cannot drift from the sources;
  * the oracle must reproduce every fixture (always runs; the GPU box has the .npz but no reference);
  * oracle/minifst.py, the stand-in for the OpenFst calls, is checked on hand-computed machines.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_reference_fixtures as MF  # noqa: E402
from conftest import GOLDEN, epoch_config, halfphone_config  # noqa: E402
from oracle import minifst, ref_exec as R, snickery_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "reference_exec.npz"))


@pytest.fixture(scope="module")
def inputs(fx):
    return {k[3:]: fx[k] for k in fx.files if k.startswith("in_")}


# ------------------------------------------------------------------------------------ fixtures == reference, now
@pytest.mark.skipif(not R.available(), reason="reference sources not present (GPU box)")
def test_committed_fixtures_equal_a_fresh_run_of_the_reference(fx, golden_epoch, golden_halfphone):
    inp = MF.input_arrays(golden_epoch, golden_halfphone)
    for k, v in inp.items():
        assert np.array_equal(fx["in_" + k], v), "seeded input %s changed" % k
    fresh = MF.generate(golden_epoch, golden_halfphone, inp)
    assert sorted(fresh) == sorted(k for k in fx.files if not k.startswith("in_"))
    for k, v in fresh.items():
        v = np.asarray(v)
        assert np.array_equal(fx[k], v, equal_nan=v.dtype.kind == "f"), "fixture %s is stale" % k


@pytest.mark.skipif(not R.available(), reason="reference sources not present (GPU box)")
def test_transform_is_syntactic_only():
    """The Python 2 -> 3 transform rewrites statements, never expressions: spot-check its pieces."""
    src = ("def f(self, (a, b), w=4):\n"
           "    print 'x', a,\n"
           "    print >> self, a / b\n"
           "    if a: print\n"
           "    raise ValueError, \"bad \" \\\n"
           "                      \"value\"\n")
    ns = dict(R._py2_builtins())
    exec(R.py2to3(src), ns)

    class Sink:
        def __init__(self):
            self.s = ""

        def write(self, x):
            self.s += x
    sink = Sink()
    with pytest.raises(ValueError, match="bad value"):
        ns["f"](sink, (7, 2))
    assert sink.s == "3\n"                       # Python 2: 7 / 2 == 3
    assert R._py2_div(7.0, 2) == 3.5 and R._py2_div(np.int64(7), 2) == 3
    # the reference's own greedy search divides a shape by multiepoch (synth_simple.py:477): must stay an int
    seg = R.load_module("segmentaxis").segment_axis
    assert seg(np.arange(10), 4, overlap=2).tolist() == [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 6, 7], [6, 7, 8, 9]]


@pytest.mark.skipif(not R.available(), reason="reference sources not present (GPU box)")
def test_reference_quirk_monophone_tree_with_too_few_units_raises(golden_halfphone):
    """synth_halfphone.py:1384-1385 maps scipy's "missing" index n through the converter: IndexError when a phone
    has fewer than n_candidates units.  The oracle / the CUDA host pad with -1 / 1e15 instead (the pre-filled
    arrays of :1377-1378 show the intent); this test documents the difference."""
    gh = golden_halfphone
    names = MF.halfphone_names(gh["phones"].tolist())
    names[7] = "px/px/rare_L/px/px"
    cfg = halfphone_config(n_candidates=6, preselection="monophone_then_acoustic")
    h = R.RefHalfphone(cfg, gh["F"], gh["Jc"], train_unit_names=names)
    with pytest.raises(IndexError):
        h.preselect_units_monophone_then_acoustic(gh["targets"][:1], [names[7]])


# ------------------------------------------------------------------------------------ oracle == fixtures
def test_oracle_greedy_matches_reference(fx, golden_epoch):
    ge = golden_epoch
    for tag, cfg in MF.epoch_cases().items():
        o = O.OracleSynthesiser(cfg, ge["F"], ge["Jc"])
        if "truncate_target_streams" in cfg:
            o.truncate_target_streams(cfg["truncate_target_streams"])
            o.truncate_join_streams(cfg["truncate_join_streams"])
        o.get_tree_for_greedy_search()
        assert np.array_equal(o.target_weight_vector, fx["%s_wt" % tag])
        for i in MF.epoch_targets(tag):
            uf = MF.epoch_unit_features(cfg, ge, i, o.target_weight_vector, getattr(o, "target_truncation_vector", None))
            assert o.greedy_joint_search(uf) == fx["%s_path_%d" % (tag, i)].tolist(), (tag, i)
            assert o.greedy_joint_search(uf, engine="brute") == fx["%s_path_%d" % (tag, i)].tolist(), (tag, i)
    # the older golden file (generated by the oracle) agrees with the reference run on the same inputs
    for i in range(3):
        assert np.array_equal(ge["path_%d" % i], fx["cfg1_m6_path_%d" % i])
    assert np.array_equal(fx["identity_path"], np.arange(400, 400 + 72, 6))
    assert np.array_equal(ge["identity_path"], fx["identity_path"])


def test_oracle_halfphone_epoch_layout_and_scores_match_reference(fx, golden_epoch):
    ge = golden_epoch
    for m in (1, 3):
        cfg = dict(epoch_config(multiepoch=m), halfphone_epoch_join_layout=True)
        o = O.OracleSynthesiser(cfg, ge["F"], MF.hp_epoch_join(ge["Jc"]))
        o.get_tree_for_greedy_search()
        uf = MF.epoch_unit_features(cfg, ge, 1, o.target_weight_vector, None)[:60]
        p = o.greedy_joint_search(uf)
        assert p == fx["hpepoch_m%d_path" % m].tolist()
        assert np.array_equal(o.get_target_scores_per_stream(o.window_targets(uf), p), fx["hpepoch_m%d_tscores" % m])
        assert np.array_equal(o.get_join_scores_per_stream(p), fx["hpepoch_m%d_jscores" % m])


def test_oracle_acoustic_preselection_and_viterbi_match_reference(fx, golden_halfphone):
    gh = golden_halfphone
    for K in (12, 50):
        o = O.OracleSynthesiser(halfphone_config(n_candidates=K), gh["F"], gh["Jc"])
        o.build_acoustic_tree()
        cand, dist = o.preselect_units_acoustic(gh["targets"])
        assert np.array_equal(cand, fx["hp_k%d_cand" % K]) and np.array_equal(dist, fx["hp_k%d_dist" % K])
        p64, c64 = o.viterbi_search(cand, dist, return_cost=True)
        p32, c32 = o.viterbi_search(cand, dist, arithmetic="openfst32", return_cost=True)
        pnp, cnp = O.viterbi_search_numpy(o, cand, dist, return_cost=True)
        ref = fx["hp_k%d_path" % K].tolist()
        assert p64 == ref and p32 == ref and pnp == ref
        assert fx["hp_k%d_path_py2str" % K].tolist() == ref
        # float32 accumulation in OpenFst's arc order: the oracle's openfst32 mode is bit-identical to the stand-in
        assert np.float32(c32) == fx["hp_k%d_cost_py2str" % K]
        assert abs(c64 - float(fx["hp_k%d_cost" % K])) <= 1e-6 * c64
    assert np.array_equal(gh["knn_idx"], fx["hp_k12_cand"]) and np.array_equal(gh["vit_path_f64"], fx["hp_k12_path"])


def test_oracle_lattice_semantics_match_reference(fx, inputs, golden_halfphone):
    gh = golden_halfphone
    o = O.OracleSynthesiser(halfphone_config(n_candidates=12), gh["F"], gh["Jc"])
    for name in MF.lattice_names():
        cand, dist = inputs[name + "_cand"], inputs[name + "_dist"]
        ref_path, ref_cost = fx[name + "_path"].tolist(), float(fx[name + "_cost"])
        p, c = o.viterbi_search(cand, dist, return_cost=True)
        pn, cn = O.viterbi_search_numpy(o, cand, dist, return_cost=True)
        assert p == ref_path and pn == ref_path, name
        if ref_path:
            assert abs(c - ref_cost) <= 2e-6 * ref_cost and abs(cn - ref_cost) <= 2e-6 * ref_cost
        else:
            assert c == np.inf
    assert fx["lat_blocked_path"].size == 0                              # a fully padded frame: no path, []
    assert np.array_equal(fx["lat_natural_path"], inputs["lat_natural_cand"][:, 0])   # natural joins cost exactly 0
    a, b = inputs["lat_natural_cand"][0], inputs["lat_natural_cand"][1]
    tile = np.array([[o.join_cost_cache(inputs["lat_natural_cand"][:2]).get((int(x), int(y)), np.inf) for y in b] for x in a])
    assert np.array_equal(tile, fx["lat_natural_tile0"]) and tile[0, 0] == 0.0


def test_oracle_label_preselection_matches_reference(fx, golden_halfphone):
    gh = golden_halfphone
    names = MF.halfphone_names(gh["phones"].tolist())
    tnames = [names[i] for i in (100, 101, 102, 300, 301, 302, 640, 641)]
    unit_index = {}
    for i, q in enumerate(names):
        f = q.split("/")
        mono = f[2]
        forms = (mono, "/".join(f[1:3]) if mono.endswith("_L") else "/".join(f[2:4]), "/".join(f[1:4]), q)
        for form in forms:
            unit_index.setdefault(form, []).append(i)
    o = O.OracleSynthesiser(halfphone_config(n_candidates=12), gh["F"], gh["Jc"])
    cq, dq = o.preselect_units_quinphone(gh["targets"][:8], tnames, unit_index)
    assert np.array_equal(cq, fx["quin_cand"]) and np.array_equal(dq, fx["quin_dist"])
    assert o.viterbi_search(cq, dq) == fx["quin_path"].tolist()
    om = O.OracleSynthesiser(halfphone_config(n_candidates=6, preselection="monophone_then_acoustic"), gh["F"], gh["Jc"])
    om.build_phonetrees(names)
    cm, dm = om.preselect_units_monophone_then_acoustic(gh["targets"][:8], tnames)
    assert np.array_equal(cm, fx["mono_cand"]) and np.array_equal(dm, fx["mono_dist"])


def test_oracle_numpy_helpers_match_reference(fx, inputs):
    assert np.array_equal(O.segment_axis0(inputs["seg_a"], 6, 5), fx["seg_6_5"])
    assert np.array_equal(O.segment_axis0(inputs["seg_a"], 6, 0), fx["seg_6_0"]) and fx["seg_6_0"].shape == (3, 6, 3)
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        st = O.standardise(np.array(inputs["std_speech"]), inputs["std_mean"].astype(cast), inputs["std_std"].astype(cast))
        assert st.dtype == fx["std_out_" + nm].dtype and np.array_equal(st, fx["std_out_" + nm])
        assert np.array_equal(O.weight(st, np.linspace(0.1, 1.0, 61)), fx["std_weighted_" + nm])
    assert np.array_equal(O.taper_matrix(O.zero_pad_matrix(np.array(inputs["taper_frag"]), 2, 0), 4), fx["taper_out"])
    t32 = O.taper_matrix(np.array(inputs["taper_frag"]), 4)
    assert t32.dtype == np.float32 and np.array_equal(t32, fx["taper_out_f32"])


# ------------------------------------------------------------------------------------ the OpenFst stand-in
def _compile(lines):
    c = minifst.Compiler()
    for l in lines:
        print(l, file=c)
    return c.compile()


def test_minifst_known_answers():
    # T: two frames, labels 1/2 then 3; J: 1->3 costs 5, 2->3 costs 1, with epsilon entry and labelled exit arcs
    T = _compile(["0 1 1 1 0.5", "0 1 2 2 2.25", "1 2 3 3 0.125", "2"])
    T.arcsort(st="olabel")
    J = _compile(["0 1 0 0", "0 2 0 0", "0 3 0 0", "1 3 1 1 5", "2 3 2 2 1", "1 4 1 1", "2 4 2 2", "3 4 3 3", "4"])
    C = minifst.compose(T, J)
    sp = minifst.shortestpath(C)
    # best: 2 (2.25) + join 1 + 3 (0.125) = 3.375 beats 1 (0.5) + 5 + 0.125
    assert float(sp.path_weight) == 3.375
    rows = [l.split("\t") for l in sp.text().split("\n") if l]
    arcs = sorted((int(r[0]), int(r[2])) for r in rows if len(r) in (4, 5))
    assert [lab for _, lab in reversed(arcs) if lab != 0] == [2, 3]
    assert [r for r in rows if len(r) == 1] == [["0"]]           # the path's final state is state 0 (built backwards)
    assert int(rows[0][0]) == len(arcs)                          # ... and the listing opens with the start state, the highest id
    # a single frame: J is empty (no pairs), the composition has no successful path
    T1 = _compile(["0 1 1 1 0.5", "1"])
    J0 = _compile(["0"])
    assert minifst.shortestpath(minifst.compose(T1, J0)).text() == ""
    # float32 Times in arc order: (d + (D + J)), not ((d + D) + J)
    big, tiny = np.float32(16777216.0), np.float32(1.0)
    Tf = _compile(["0 1 1 1 1", "1 2 2 2 1", "2"])
    Jf = _compile(["0 1 0 0", "0 2 0 0", "1 2 1 1 %r" % float(big), "1 3 1 1", "2 3 2 2", "3"])
    spf = minifst.shortestpath(minifst.compose(Tf, Jf))
    assert spf.path_weight == np.float32(np.float32(np.float32(0) + np.float32(tiny + big)) + np.float32(tiny + 0))


def test_minifst_py2_str_mode_rounds_through_twelve_digits():
    w = 0.1234567890123456
    minifst.Compiler.py2_str = True
    try:
        f = _compile(["0 1 1 1 %s" % w, "1"])
    finally:
        minifst.Compiler.py2_str = False
    assert f.arcs[0][0][2] == np.float32(float("%.12g" % w))


def test_oracle_and_host_label_logic_match_reference_halfphone_stats(fx, inputs):
    """train_halfphone.get_halfphone_stats as the reference runs it: the oracle's restatement and the product's host-side
    frame picking (snickery_b200.synth.halfphone_unit_points) give the same names, timings and sampled frames."""
    from snickery_b200.synth import halfphone_unit_points
    labs = MF.hp3_labels(inputs["hp3_state_ends"])
    names, starts, middles, ends = halfphone_unit_points(labs, 70)
    assert names.tolist() == fx["hp3_names"].tolist()
    assert np.array_equal(np.stack([starts, ends], axis=1), fx["hp3_timings"])
    assert ends.max() == 69                                      # the last state ran past the utterance: clipped
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        st = O.standardise(np.array(inputs["hp3_speech"]), inputs["std_mean"].astype(cast), inputs["std_std"].astype(cast))
        for rep, npts in (("onepoint", 1), ("twopoint", 2), ("threepoint", 3)):
            n2, feats, timings = O.halfphone_stats(st, labs, rep)
            assert n2.tolist() == names.tolist() and timings == list(zip(starts.tolist(), ends.tolist()))
            assert np.array_equal(O.weight(feats, np.linspace(0.2, 1.1, 61 * npts)), fx["hp3_%s_%s" % (rep, nm)])
            pts = {"onepoint": [middles], "twopoint": [starts, ends], "threepoint": [starts, middles, ends]}[rep]
            assert np.array_equal(np.hstack([st[p] for p in pts]), feats)
