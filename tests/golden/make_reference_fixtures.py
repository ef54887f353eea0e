"""Fixtures produced by EXECUTING THE REFERENCE'S OWN CODE (oracle/ref_exec.py) on seeded inputs.

    python tests/golden/make_reference_fixtures.py          # needs /root/reference; writes reference_exec.npz

Every expected output in `reference_exec.npz` comes out of functions lifted from /root/reference/script
(synth_simple.py, synth_halfphone.py, fst_functions_wrapped.py, segmentaxis.py, speech_manip.py,
data_manipulation.py, matrix_operations.py) after a mechanical Python 2 -> 3 transform; the k-NN engine is
the reference's own scipy cKDTree and the OpenFst calls land in oracle/minifst.py.  Inputs that are not already
in epoch_greedy.npz / halfphone_viterbi.npz (the databases) are stored beside the outputs.

`cases()` is shared with tests/test_reference_exec.py, which regenerates everything in memory when
/root/reference is present and compares it bit for bit with the committed file, and with the GPU parity tests,
which compare the CUDA path with the committed file (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import epoch_config, halfphone_config  # noqa: E402

OUT = os.path.join(HERE, "reference_exec.npz")


def halfphone_names(phones):
    """Synthetic internal quinphone labels 'll/l/c_X/r/rr' (const.py:5, label_manip.py:16-32)."""
    names = []
    for i, p in enumerate(phones):
        side = "_L" if i % 2 == 0 else "_R"
        l, r = phones[max(i - 1, 0)], phones[min(i + 1, len(phones) - 1)]
        names.append("p%d/p%d/p%d%s/p%d/p%d" % (phones[max(i - 2, 0)], l, p, side, r, phones[min(i + 2, len(phones) - 1)]))
    return names


def hp_epoch_join(Jc1):
    """Two-frame join windows as train_halfphone.py writes them for epoch voices: [N+1, 302] = frame u-1 | frame u."""
    return np.ascontiguousarray(np.hstack([Jc1, np.vstack([Jc1[1:], Jc1[-1:]])]))


def random_lattice(rng, n_units, T, K, kind):
    """Candidate lattices with the quirks the reference's lattice code reacts to (synth_halfphone.py:3238-3268)."""
    cand = rng.integers(1, n_units - 1, size=(T, K)).astype(np.int64)
    if kind == "quirks":
        cand[rng.random((T, K)) < 0.15] = -1                  # padding
        cand[rng.random((T, K)) < 0.03] = 0                   # unit 0: never joinable
        cand[rng.random((T, K)) < 0.03] = n_units - 1         # last unit: never joinable
        for t in range(T):                                    # back-off duplicates
            if K > 2 and rng.random() < 0.5:
                cand[t, K - 1] = cand[t, 0]
    elif kind == "natural":
        # consecutive units: natural joins (cost exactly 0) compete with the rest
        base = int(rng.integers(5, n_units - T - 5))
        cand[:, 0] = base + np.arange(T)
    elif kind == "blocked":
        cand[T // 2, :] = -1
    dist = rng.random((T, K)) * 3.0
    if kind == "natural":
        dist[:, 0] *= 0.05           # ... and win: the expected path is the natural one
    return cand, dist


def input_arrays(ge, gh):
    """Seeded inputs that are not part of the older golden files."""
    rng = np.random.default_rng(20261017)
    inp = {}
    n_hp = gh["F"].shape[0]
    for name, (T, K, kind) in {"lat_quirks": (14, 9, "quirks"), "lat_natural": (12, 5, "natural"),
                               "lat_blocked": (8, 6, "blocked"), "lat_k1": (9, 1, "plain"),
                               "lat_k30": (20, 30, "quirks"), "lat_k64": (10, 64, "plain")}.items():
        c, d = random_lattice(rng, n_hp, T, K, kind)
        inp[name + "_cand"], inp[name + "_dist"] = c, d
    inp["std_speech"] = rng.standard_normal((12, 61)).astype(np.float32) * 3 + 1
    inp["std_speech"][rng.random((12, 61)) < 0.1] = -1000.0
    inp["std_speech"][:, 60][::3] = -1000.0
    inp["std_mean"] = rng.standard_normal(61)
    inp["std_std"] = rng.uniform(0.5, 2.0, size=(1, 61))
    inp["taper_frag"] = rng.standard_normal((10, 7)).astype(np.float32)
    inp["seg_a"] = rng.standard_normal((23, 3))
    # half-phone target preparation (train_halfphone.py:959-1070): 6 phones x 5 states over a 70-frame utterance whose last
    # state runs past the end (clipped), un-normalised speech with unvoiced markers, normalised durations
    bounds = np.sort(rng.choice(np.arange(1, 72), size=29, replace=False))
    inp["hp3_state_ends"] = np.concatenate([bounds, [74]]).astype(np.int64)          # end frame of each of the 30 states
    inp["hp3_speech"] = rng.standard_normal((70, 61)).astype(np.float32) * 2 + 0.5
    inp["hp3_speech"][rng.random((70, 61)) < 0.05] = -1000.0
    inp["hp3_speech"][:, 60][::4] = -1000.0
    inp["hp3_dur"] = rng.standard_normal((12, 1))
    return inp


def hp3_labels(state_ends):
    """The reference's label structure: [((start, end), [ll, l, c, r, rr, state]), ...] (label_manip.py / read_label)."""
    labs, s = [], 0
    for i, e in enumerate(state_ends.tolist()):
        ph = i // 5
        labs.append(((s, int(e)), ["p%d" % (ph - 2), "p%d" % (ph - 1), "p%d" % ph, "p%d" % (ph + 1), "p%d" % (ph + 2), str(2 + i % 5)]))
        s = int(e) + 1
    return labs


def generate(ge, gh, inp):
    """Runs the reference.  Returns {name: array}."""
    from oracle import minifst, ref_exec as R
    out = {}
    weight = R.load_module("speech_manip").weight

    # ---------------- epoch voices: synth_simple.Synthesiser
    for tag, cfg in epoch_cases().items():
        r = R.RefSimple(cfg, ge["F"], ge["Jc"])
        r.get_tree_for_greedy_search()
        out["%s_wt" % tag] = np.asarray(r.target_weight_vector, dtype=np.float64)
        for i in epoch_targets(tag):
            uf = epoch_unit_features(cfg, ge, i, r.target_weight_vector, getattr(r, "target_truncation_vector", None))
            out["%s_path_%d" % (tag, i)] = np.array(r.greedy_joint_search(uf), dtype=np.int64)
    # start_state (synth_simple.py:467-470): natural continuation from unit 400 (m = 6)
    cfg = epoch_cases()["cfg1_m6"]
    r = R.RefSimple(cfg, ge["F"], ge["Jc"])
    tf = np.array(r.train_unit_features[400:400 + 6 * 12])        # unweighted-window copy BEFORE the tree reshapes it
    r.get_tree_for_greedy_search()
    out["identity_path"] = np.array(r.greedy_joint_search(tf, start_state=400), dtype=np.int64)

    # ---------------- epoch voice through synth_halfphone.Synthesiser (two-frame join windows, greedy)
    for m in (1, 3):
        cfg = dict(epoch_config(multiepoch=m), halfphone_epoch_join_layout=True)
        h = R.RefHalfphone(cfg, ge["F"], hp_epoch_join(ge["Jc"]))
        uf = epoch_unit_features(cfg, ge, 1, h.target_weight_vector, None)[:60]
        p = h.greedy_joint_search(uf)
        out["hpepoch_m%d_path" % m] = np.array(p, dtype=np.int64)
        # per-stream cost report (synth_halfphone.py:1964-1981, 2977-3008): squared errors of the chosen rows against
        # the windowed targets; stream widths are the unwindowed ones, so for m > 1 only the first frame is reported
        win = uf[: (uf.shape[0] // m) * m].reshape(-1, m * uf.shape[1])
        out["hpepoch_m%d_tscores" % m] = h.get_target_scores_per_stream(win, p)
        out["hpepoch_m%d_jscores" % m] = h.get_join_scores_per_stream(p)

    # ---------------- halfphone voice: synth_halfphone.Synthesiser + fst_functions_wrapped + minifst
    names = halfphone_names(gh["phones"].tolist())
    for K in (12, 50):
        cfg = halfphone_config(n_candidates=K)
        h = R.RefHalfphone(cfg, gh["F"], gh["Jc"], train_unit_names=names)
        cand, dist = h.preselect_units_acoustic(gh["targets"])
        out["hp_k%d_cand" % K], out["hp_k%d_dist" % K] = np.asarray(cand, dtype=np.int64), dist
        for py2 in (False, True):
            minifst.Compiler.py2_str = py2
            p = h.viterbi_search(cand, dist)
            sfx = "_py2str" if py2 else ""
            out["hp_k%d_path%s" % (K, sfx)] = np.array(p, dtype=np.int64)
            out["hp_k%d_cost%s" % (K, sfx)] = np.float32(R.last_viterbi_cost())
        minifst.Compiler.py2_str = False
    cfg = halfphone_config(n_candidates=12)
    h = R.RefHalfphone(cfg, gh["F"], gh["Jc"], train_unit_names=names)
    # join costs as the J acceptor carries them: the cost_cache dict -> tile (synth_halfphone.py:3206-3301)
    for name in lattice_names():
        cand, dist = inp[name + "_cand"], inp[name + "_dist"]
        p = h.viterbi_search(cand, dist)
        out[name + "_path"] = np.array(p, dtype=np.int64)
        out[name + "_cost"] = np.float32(R.last_viterbi_cost() if p else np.inf)
    first, second = np.meshgrid(inp["lat_natural_cand"][0], inp["lat_natural_cand"][1], indexing="ij")
    out["lat_natural_tile0"] = h.get_natural_distance_vectorised(first.ravel(), second.ravel(), order=1).reshape(first.shape)
    # quinphone preselection + distances (synth_halfphone.py:1305-1354), then the search
    tnames = [names[i] for i in (100, 101, 102, 300, 301, 302, 640, 641)]
    cq, dq = h.preselect_units_quinphone(gh["targets"][:8], tnames)
    out["quin_cand"], out["quin_dist"] = np.asarray(cq, dtype=np.int64), dq
    out["quin_path"] = np.array(h.viterbi_search(cq, dq), dtype=np.int64)
    # monophone-then-acoustic (synth_halfphone.py:1369-1396); every phone of the synthetic voice has >= 6 units
    cfgm = halfphone_config(n_candidates=6, preselection="monophone_then_acoustic")
    hm = R.RefHalfphone(cfgm, gh["F"], gh["Jc"], train_unit_names=names)
    cm, dm = hm.preselect_units_monophone_then_acoustic(gh["targets"][:8], tnames)
    out["mono_cand"], out["mono_dist"] = np.asarray(cm, dtype=np.int64), dm

    # ---------------- plain numpy helpers
    seg = R.load_module("segmentaxis").segment_axis
    out["seg_6_5"] = np.array(seg(inp["seg_a"], 6, overlap=5, axis=0))
    out["seg_6_0"] = np.array(seg(inp["seg_a"], 6, overlap=0, axis=0))          # remainder (23 % 6) is cut
    dm_ = R.load_module("data_manipulation")
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        sp = np.array(inp["std_speech"])
        st = dm_.standardise(sp, inp["std_mean"].astype(cast), inp["std_std"].astype(cast))
        out["std_out_" + nm] = st
        out["std_weighted_" + nm] = weight(st, np.linspace(0.1, 1.0, 61))
    # half-phone targets: standardise -> get_halfphone_stats -> hstack(durations) -> weight (synth_halfphone.py:1510-1548)
    th = R.load_module("train_halfphone")
    labs = hp3_labels(inp["hp3_state_ends"])
    for nm, cast in (("f64", np.float64), ("f32", np.float32)):
        st = dm_.standardise(np.array(inp["hp3_speech"]), inp["std_mean"].astype(cast), inp["std_std"].astype(cast))
        for rep, npts in (("onepoint", 1), ("twopoint", 2), ("threepoint", 3)):
            names, feats, timings = th.get_halfphone_stats(st, labs, representation_type=rep)
            out["hp3_%s_%s" % (rep, nm)] = weight(feats, np.linspace(0.2, 1.1, 61 * npts))
        out["hp3_names"] = np.array([str(n) for n in names])
        out["hp3_timings"] = np.array(timings, dtype=np.int64)
        out["hp3_threepoint_dur_" + nm] = weight(np.hstack([feats, inp["hp3_dur"]]), np.linspace(0.2, 1.1, 184))
    mo = R.load_module("matrix_operations")
    out["taper_out"] = mo.taper_matrix(mo.zero_pad_matrix(np.array(inp["taper_frag"]), 2, 0), 4)
    out["taper_out_f32"] = mo.taper_matrix(np.array(inp["taper_frag"]), 4)
    return out


def lattice_names():
    return ["lat_quirks", "lat_natural", "lat_blocked", "lat_k1", "lat_k30", "lat_k64"]


def epoch_cases():
    return {
        "cfg1_m6": epoch_config(),                                                     # slt_simplified_mini.cfg weights
        "is2018_m6": epoch_config(multiepoch=6, tsw=(0.5, 0.5)),                        # IS2018_nick_simplified.cfg weights
        "cfg1_m1": epoch_config(multiepoch=1),
        "cfg1_m3": epoch_config(multiepoch=3),
        "cfg1_m4": epoch_config(multiepoch=4),
        "trunc_m2": dict(epoch_config(multiepoch=2), truncate_target_streams=[20, -1], truncate_join_streams=[30, 10, 0, 1]),
    }


def epoch_targets(tag):
    return (0, 1, 2) if tag == "cfg1_m6" else (1,)


def epoch_unit_features(cfg, ge, i, target_weight_vector, truncation):
    """What synth_utt hands to greedy_joint_search (synth_simple.py:381-397): weighted (and truncated) frames.
    The golden targets were stored weighted with the config-1 vector; undo that to get the standardised frames."""
    w1 = np.array([0.1 * 0.8] * 60 + [1.0 * 0.8])
    x = (ge["targets_%d" % i] / w1).astype(np.float32)          # float32, as read from a feature file
    uf = x * np.asarray(target_weight_vector, dtype=np.float64).reshape(1, -1)
    if truncation is not None:
        uf = uf[:, truncation]
    return uf


def main():
    from oracle import ref_exec as R
    assert R.available(), "needs the reference sources under %s" % R.SCRIPT
    ge = np.load(os.path.join(HERE, "epoch_greedy.npz"))
    gh = np.load(os.path.join(HERE, "halfphone_viterbi.npz"))
    inp = input_arrays(ge, gh)
    out = generate(ge, gh, inp)
    blob = {"in_" + k: v for k, v in inp.items()}
    blob.update(out)
    np.savez_compressed(OUT, **blob)
    print(OUT, os.path.getsize(OUT), "bytes,", len(out), "reference outputs")


if __name__ == "__main__":
    main()
