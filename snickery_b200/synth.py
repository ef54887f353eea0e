"""Python host mirroring the search-relevant interface of snickery's Synthesiser classes.

Same method names, config keys and argument meaning as the reference
(script/synth_simple.py:48-503, script/synth_halfphone.py:151-1436) for the hot path only:

    set_join_weights / set_target_weights / get_tree_for_greedy_search / reconfigure_settings
    greedy_joint_search / preselect_units_acoustic / viterbi_search
    get_target_scores_per_stream / get_join_scores_per_stream

Everything numeric runs on the GPU through the C ABI (engine.py); waveform generation,
HDF5 voices and label handling are out of scope.  Arrays that the
reference reads from its HDF5 voice are passed to the constructor instead.
Config keys kept: target_stream_weights, join_stream_weights, join_cost_weight, multiepoch,
search_epsilon, n_candidates, greedy_search, target_representation, add_duration_as_target,
duration_target_weight, stream_list_*, datadims_*.
"""
from __future__ import annotations

import timeit

import numpy as np

from . import engine
from .kdtree import GpuKDTree

APPLY_JCW_ON_TOP = True  # reference script/synth_simple.py:44
VERY_BIG_WEIGHT_VALUE = 1000000000000000.0  # const.py:3
LABEL_DELIMITER = "/"  # const.py:5


def break_quinphone(quinphone):
    """label_manip.py:16-32: internal 'a/b/c_L/d/e' label -> (mono, diphone, triphone, quinphone)."""
    q = quinphone.split(LABEL_DELIMITER)
    assert len(q) == 5
    mono = q[2]
    tri = LABEL_DELIMITER.join(q[1:4])
    if mono.endswith("_L"):
        di = LABEL_DELIMITER.join(q[1:3])
    elif mono.endswith("_R"):
        di = LABEL_DELIMITER.join(q[2:4])
    else:
        raise ValueError("halfphone label %r lacks the _L/_R suffix" % mono)
    return (mono, di, tri, quinphone)
TARGET_REP_WIDTHS = {"onepoint": 1, "twopoint": 2, "threepoint": 3, "epoch": 1, "sample": 1}  # const.py:17


def read_config(config):
    """A snickery config is Python source exec'd into a dict (synth_simple.py:56-58)."""
    if isinstance(config, dict):
        return dict(config)
    out = {}
    with open(config) as f:
        exec(compile(f.read(), config, "exec"), out)
    out.pop("__builtins__", None)
    return out


def segment_axis_cut(a, length):
    """Non-overlapping windows along axis 0, remainder cut (segmentaxis.py:76-77,94-96)."""
    n = a.shape[0] // length
    if n == 0:
        raise ValueError("Not enough data points to segment array in 'cut' mode; try 'pad' or 'wrap'")
    return a[: n * length].reshape(n, length * a.shape[1])


def halfphone_unit_points(labels, n_frames):
    """Frame indices that describe every half-phone of an utterance (train_halfphone.py:959-1070).  labels: the
    reference's 5-state alignment, a list of ((start, end), lab) with lab = [ll, l, c, r, rr, state] and states '2'..'6'
    per phone.  States 2-3 make the left half-phone (first frame: start of 2, middle: end of 2, last: end of 3), states
    4-6 the right one (start of 4, end of 5, end of 6); ends past the utterance are clipped to its last frame.
    Returns (names, starts, middles, ends): names 'll/l/c_L/r/rr' style, int64 arrays of equal length."""
    if len(labels) % 5 != 0:
        raise AssertionError("There must be 5 states for each phone in label")
    names, starts, middles, ends = [], [], [], []
    for (s, e), lab in labels:
        e = min(e, n_frames - 1)
        if len(lab) != 6:
            raise AssertionError("label must be a quinphone plus state")
        state = lab[-1]
        if state in ("2", "4"):
            quin = list(lab[:5])
            quin[2] += "_L" if state == "2" else "_R"
            if any(LABEL_DELIMITER in part for part in quin):
                raise AssertionError("delimiter %s occurs in one or more name element (%s)" % (LABEL_DELIMITER, quin))
            names.append(LABEL_DELIMITER.join(quin))
            starts.append(s)
            if state == "2":
                middles.append(e)
        elif state == "5":
            middles.append(e)
        elif state in ("3", "6"):
            ends.append(e)
        else:
            raise ValueError("bad state number")
    if not (len(names) == len(starts) == len(middles) == len(ends) == 2 * (len(labels) // 5)):
        raise AssertionError("state alignment does not describe two half-phones per phone")
    as_int = lambda v: np.asarray(v, dtype=np.int64)
    return np.array(names), as_int(starts), as_int(middles), as_int(ends)


class Synthesiser:
    def __init__(self, config, train_unit_features_unweighted, join_contexts_unweighted, device=0, verbose=False,
                 train_unit_names=None):
        self.config = read_config(config)
        self.verbose = verbose
        self.stream_list_target = self.config["stream_list_target"]
        self.stream_list_join = self.config["stream_list_join"]
        self.datadims_target = self.config["datadims_target"]
        self.datadims_join = self.config["datadims_join"]
        self.target_representation = self.config.get("target_representation", "epoch")
        F = np.ascontiguousarray(train_unit_features_unweighted, dtype=np.float32)
        Jc = np.ascontiguousarray(join_contexts_unweighted, dtype=np.float32)
        self.number_of_units = F.shape[0]
        greedy_epoch = self.config.get("greedy_search", False) and self.target_representation == "epoch"
        self._multiepoch = int(self.config.get("multiepoch", 1)) if greedy_epoch else 1
        layout = engine.LAYOUT_SIMPLE
        if self.config.get("halfphone_epoch_join_layout", False):   # synth_halfphone.py:552-553
            layout = engine.LAYOUT_HALFPHONE_EPOCH
        self.db = engine.UnitDatabase(F, Jc, multiepoch=self._multiepoch, layout=layout, device=device)
        self._wt = self._wj = None
        self._dirty = True
        self._tmask = self._jmask = None          # truncate_*_streams column masks
        self._F_raw = F
        self.train_unit_names = train_unit_names
        self.phonetrees = None
        jcw = self.config["join_cost_weight"]
        if APPLY_JCW_ON_TOP:   # synth_simple.py:128-133
            self.set_target_weights(np.array(self.config["target_stream_weights"]) * (1.0 - jcw))
            self.set_join_weights(np.array(self.config["join_stream_weights"]) * jcw)
        else:
            self.set_target_weights(self.config["target_stream_weights"])
            self.set_join_weights(self.config["join_stream_weights"])
        if "truncate_target_streams" in self.config:   # synth_simple.py:136-139
            self.truncate_target_streams(self.config["truncate_target_streams"])
        if "truncate_join_streams" in self.config:
            self.truncate_join_streams(self.config["truncate_join_streams"])
        self._push_weights()   # the reference weights its arrays in the constructor (synth_simple.py:128-133)
        if train_unit_names is not None and self.target_representation != "epoch":
            # label index for quinphone preselection (synth_halfphone.py:281-292)
            self.unit_index = {}
            for i, quinphone in enumerate(train_unit_names):
                for form in break_quinphone(quinphone):
                    self.unit_index.setdefault(form, []).append(i)
            if self.config.get("preselection_method") == "monophone_then_acoustic":
                self._build_phonetrees()
        if greedy_epoch:
            self.get_tree_for_greedy_search()
        elif self.config.get("preselection_method", "quinphone") == "acoustic":
            self.tree = GpuKDTree(None, _db=self.db, _space=engine.SPACE_TARGET)

    # ---- clocks, as the reference prints them (synth_simple.py:760-768)
    @classmethod
    def from_voice(cls, config, datafile, device=0, verbose=False):
        """Load the reference's database dump (synth_simple.py:76-106) and build the resident
        database from it.  The file's float32 mean_target / std_target become the device-side
        standardisation, so un-normalised test speech can go straight to greedy_joint_search_unnorm_batch."""
        from .hdf5_voice import load_voice
        voice = load_voice(datafile)
        names = voice.get("train_unit_names")
        if names is not None:
            names = [n.decode("ascii", "replace") if isinstance(n, bytes) else str(n) for n in names.tolist()]
        self = cls(config, voice["train_unit_features"], voice["join_contexts"], device=device, verbose=verbose,
                   train_unit_names=names)
        self.train_filenames = voice.get("filenames")
        self.unit_index_within_sentence = voice.get("unit_index_within_sentence_dset")
        self.mean_vec_join, self.std_vec_join = voice["mean_join"], voice["std_join"]
        self.set_standardisation(voice["mean_target"], voice["std_target"])
        return self

    def start_clock(self, comment):
        if self.verbose:
            print("%s... " % comment, end="")
        return (timeit.default_timer(), comment)

    def stop_clock(self, start, width=40):
        t0, comment = start
        if self.verbose:
            print("%s--> took %.2f seconds" % ((width - len(comment)) * " ", timeit.default_timer() - t0))

    # ---- weights (synth_simple.py:234-274; synth_halfphone.py:682-737)
    def _per_coeff(self, weights, streams, dims):
        assert len(weights) == len(streams), (weights, streams)
        vec = []
        for i, stream in enumerate(streams):
            vec.extend([float(weights[i])] * dims[stream])
        return vec

    def set_join_weights(self, weights):
        vec = self._per_coeff(weights, self.stream_list_join, self.datadims_join)
        if self.config.get("halfphone_epoch_join_layout", False):
            vec = vec + vec   # natural2 cost doubles the vector (synth_halfphone.py:693-695)
        self.join_weight_vector = np.array(vec, dtype=np.float64)
        self._wj = self.join_weight_vector
        self._dirty = True

    def set_target_weights(self, weights):
        vec = self._per_coeff(weights, self.stream_list_target, self.datadims_target)
        vec = vec * TARGET_REP_WIDTHS[self.target_representation]
        if self.config.get("add_duration_as_target", False):
            vec.append(self.config.get("duration_target_weight", 0.0))
        self.target_weight_vector = np.array(vec, dtype=np.float64)
        self._wt = self.target_weight_vector
        self._dirty = True

    def _push_weights(self):
        if self._dirty:
            t = self.start_clock("re-weight resident database")
            wt = self._wt if self._tmask is None else self._wt * self._tmask
            wj = self._wj if self._jmask is None else self._wj * self._jmask
            self.db.set_weights(wt, wj)
            if self.phonetrees:
                for tree in self.phonetrees.values():
                    tree._db.set_weights(wt, np.ones(1))
            self._dirty = False
            self.stop_clock(t)

    # ---- stream truncation (synth_simple.py:968-992).  Dropping columns on both sides of a distance is
    # the same as giving them zero weight, so the resident matrices keep their shape.
    def get_selection_vector(self, stream_list, stream_dims, truncation_values):
        assert len(truncation_values) == len(stream_list), (truncation_values, stream_list)
        selection_vector = []
        start = 0
        for stream, trunc in zip(stream_list, truncation_values):
            stream_dim = stream_dims[stream]
            if trunc == -1:
                trunc = stream_dim
            assert trunc <= stream_dim, "stream %s has only %s dims, cannot truncate to %s" % (stream, stream_dim, trunc)
            selection_vector.extend(range(start, start + trunc))
            start += stream_dim
        return selection_vector

    def truncate_join_streams(self, truncation_values):
        sel = self.get_selection_vector(self.stream_list_join, self.datadims_join, truncation_values)
        mask = np.zeros(len(self._wj))
        width = sum(self.datadims_join[s] for s in self.stream_list_join)
        for rep in range(len(self._wj) // width):
            mask[np.array(sel) + rep * width] = 1.0
        self._jmask = mask
        self._dirty = True

    def truncate_target_streams(self, truncation_values):
        sel = self.get_selection_vector(self.stream_list_target, self.datadims_target, truncation_values)
        assert TARGET_REP_WIDTHS[self.target_representation] == 1, "truncation is only used with epoch voices"
        mask = np.zeros(len(self._wt))
        mask[sel] = 1.0
        self._tmask = mask
        self.target_truncation_vector = sel
        self._dirty = True

    def _full_width(self, unit_features):
        """synth_utt hands greedy_joint_search the TRUNCATED columns (synth_simple.py:392-397); put them back
        at their places (the other columns carry zero weight)."""
        u = np.asarray(unit_features, dtype=np.float64)
        sel = getattr(self, "target_truncation_vector", None)
        if sel is not None and u.shape[1] == len(sel) and len(sel) != self.db.Dt:
            full = np.zeros((u.shape[0], self.db.Dt))
            full[:, sel] = u
            return full
        return u

    def get_tree_for_greedy_search(self):
        """synth_simple.py:190-230.  Nothing is built: re-weighting refreshes the device operands."""
        self._push_weights()
        self.joint_tree = GpuKDTree(None, _db=self.db, _space=engine.SPACE_JOINT)
        return self.joint_tree

    def reconfigure_settings(self, changed_config_values):
        """synth_simple.py:776-830: only weight-type changes touch the search structures."""
        assert "multiepoch" not in changed_config_values or \
            changed_config_values["multiepoch"] == self.config.get("multiepoch", 1), \
            "multiepoch changes the resident layout: build a new Synthesiser"
        self.config.update(changed_config_values)
        jcw = self.config["join_cost_weight"]
        if APPLY_JCW_ON_TOP:
            self.set_target_weights(np.array(self.config["target_stream_weights"]) * (1.0 - jcw))
            self.set_join_weights(np.array(self.config["join_stream_weights"]) * jcw)
        else:
            self.set_target_weights(self.config["target_stream_weights"])
            self.set_join_weights(self.config["join_stream_weights"])
        self._push_weights()

    # ---- greedy (synth_simple.py:458-503)
    def greedy_joint_search(self, unit_features, start_state=-1, holdout=[]):
        assert self.config["target_representation"] == "epoch"
        assert len(holdout) == 0, "holdout filtering is commented out in the reference (synth_simple.py:493-496)"
        t = self.start_clock("Greedy search")
        path = self.greedy_joint_search_batch([unit_features], [start_state])[0]
        self.stop_clock(t)
        return path

    def greedy_joint_search_batch(self, unit_features_list, start_states=None, return_dists=False):
        self._push_weights()
        feats = [self._full_width(u) for u in unit_features_list]
        return self.db.greedy_batch(feats, start_states, return_dists=return_dists)

    # ---- target preparation on the device (synth_simple.py:371-391)
    def set_standardisation(self, mean_vec_target, std_vec_target, special_uv_value=-1000.0, uv_scaling_factor=20.0):
        """The voice's mean_vec_target / std_vec_target (synth_simple.py:96-97) and const.py:12-14."""
        self.mean_vec_target = np.asarray(mean_vec_target)      # dtype kept: float32 statistics -> float32 arithmetic
        self.std_vec_target = np.asarray(std_vec_target)
        # a half-phone voice standardises FRAMES (dim columns) before it samples them into Dt-wide units: those
        # statistics reach the device per point in halfphone_targets()
        if self.mean_vec_target.size == self.db.Dt:
            self.db.set_standardisation(self.mean_vec_target, self.std_vec_target, special_uv_value, uv_scaling_factor)

    def prepare_targets(self, unnorm_speech):
        """weight(standardise(unnorm_speech), target_weight_vector): compose_speech's float32 output in,
        the float64 unit_features of synth_utt out (full width; truncated columns carry zero weight).
        With REPLICATE_IS2018_EXP the first and the last frame are dropped (synth_simple.py:384-387)."""
        self._push_weights()
        return self.db.prepare_targets(self._trim_speech(unnorm_speech))

    def halfphone_targets(self, unnorm_speech, labels, durations=None):
        """The half-phone unit_features of synth_utt (synth_halfphone.py:1510-1548): standardise the utterance, sample
        every half-phone at the frames its state alignment names (get_halfphone_stats, train_halfphone.py:959-1070),
        append the normalised durations (config add_duration_as_target) and weight -- the frame picking on the host
        (labels are Python objects), everything else in one device kernel on float32 input.  Needs set_standardisation
        with the FRAME statistics (synth_halfphone.py mean_vec_target / std_vec_target); returns
        (unit_names, unit_features float64 [n, Dt], unit_timings)."""
        u = np.asarray(unnorm_speech, dtype=np.float32)
        names, starts, middles, ends = halfphone_unit_points(labels, u.shape[0])
        rep = self.target_representation
        points = {"onepoint": [middles], "twopoint": [starts, ends], "threepoint": [starts, middles, ends]}[rep]
        points = np.stack(points, axis=1)
        npts, dim = points.shape[1], u.shape[1]
        extra = 1 if durations is not None else 0
        if npts * dim + extra != self.db.Dt:
            raise ValueError("%s targets of %d-dim frames%s are %d wide, the voice's are %d" %
                             (rep, dim, " + duration" if extra else "", npts * dim + extra, self.db.Dt))
        # the device applies column-wise statistics: the frame statistics once per point (the duration is only weighted)
        mean = np.concatenate([np.ravel(self.mean_vec_target)] * npts + [np.zeros(extra, self.mean_vec_target.dtype)])
        std = np.concatenate([np.ravel(self.std_vec_target)] * npts + [np.ones(extra, self.std_vec_target.dtype)])
        self._push_weights()
        self.db.set_standardisation(mean, std, -1000.0, 20.0)
        feats = self.db.halfphone_targets(u, points, durations)
        return names, feats, list(zip(starts.tolist(), ends.tolist()))

    def _trim_speech(self, unnorm_speech):
        u = np.asarray(unnorm_speech, dtype=np.float32)
        return u[1:-1, :] if self.config.get("REPLICATE_IS2018_EXP", False) else u

    def greedy_joint_search_unnorm_batch(self, unnorm_speech_list, start_states=None, return_dists=False):
        """greedy_joint_search straight from un-normalised speech: standardise + weight are fused into the
        device-side query assembly, so the host does no per-utterance numpy and uploads float32."""
        self._push_weights()
        feats = [self._trim_speech(u) for u in unnorm_speech_list]
        return self.db.greedy_batch(feats, start_states, return_dists=return_dists, unnorm=True)

    # ---- preselection (synth_halfphone.py:1359-1366, 1346-1351)
    def preselect_units_acoustic(self, unit_features):
        self._push_weights()
        t = self.start_clock("Acoustic select units ")
        distances, candidates = self.tree.query(unit_features, k=self.config["n_candidates"])
        self.stop_clock(t)
        return (candidates, distances)

    def _build_phonetrees(self):
        """One search structure per monophone (synth_halfphone.py:385-402)."""
        monophones = np.array([q.split(LABEL_DELIMITER)[2] for q in self.train_unit_names])
        self.phonetrees, self.phonetrees_index_converters = {}, {}
        wt = self._wt if self._tmask is None else self._wt * self._tmask
        for phone in dict.fromkeys(monophones.tolist()):
            sel = monophones == phone
            self.phonetrees[phone] = GpuKDTree.from_weighted(self._F_raw[sel, :], wt, device=self.db.device)
            self.phonetrees_index_converters[phone] = np.arange(self.number_of_units)[sel]

    def preselect_units_monophone_then_acoustic(self, unit_features, unit_names):
        """synth_halfphone.py:1369-1396; -1 / VERY_BIG_WEIGHT_VALUE padding where a phone has too few units.
        Targets of the same phone are searched in one batched call."""
        self._push_weights()
        K = self.config["n_candidates"]
        unit_features = np.asarray(unit_features, dtype=np.float64)
        m = unit_features.shape[0]
        candidates = np.ones((m, K), dtype=int) * -1
        distances = np.ones((m, K)) * VERY_BIG_WEIGHT_VALUE
        monophones = np.array([q.split(LABEL_DELIMITER)[2] for q in unit_names])
        assert len(monophones) == m, (len(monophones), m)
        for phone in dict.fromkeys(monophones.tolist()):
            assert phone in self.phonetrees, "unseen monophone %s" % phone
            rows = np.flatnonzero(monophones == phone)
            conv = self.phonetrees_index_converters[phone]
            kk = min(K, conv.size)
            d, i = self.phonetrees[phone].query(unit_features[rows], k=kk)
            candidates[rows[:, None], np.arange(kk)[None, :]] = conv[np.asarray(i).reshape(len(rows), kk)]
            distances[rows[:, None], np.arange(kk)[None, :]] = np.asarray(d).reshape(len(rows), kk)
        return (candidates, distances)

    def preselect_units_quinphone(self, unit_features, unit_names):
        """synth_halfphone.py:1305-1354: label-index candidates (quin -> tri -> di -> mono, duplicates kept,
        -1 padded, empty -> [1]); target distances on the GPU."""
        K = self.config["n_candidates"]
        candidates = []
        for quinphone in unit_names:
            current = []
            mono, diphone, triphone, quin = break_quinphone(quinphone)
            for form in [quin, triphone, diphone, mono]:
                for unit in self.unit_index.get(form, []):
                    current.append(unit)
                    if len(current) == K:
                        break
                if len(current) == K:
                    break
            if len(current) == 0:
                current = [1]
            current += [-1] * (K - len(current))
            candidates.append(current)
        candidates = np.array(candidates)
        distances = self.candidate_target_distances(candidates, unit_features)
        return (candidates, distances)

    def candidate_target_distances(self, candidates, unit_features):
        self._push_weights()
        return self.db.candidate_distances(candidates, unit_features)

    # ---- Viterbi (synth_halfphone.py:1399-1436)
    def viterbi_search(self, candidates, distances):
        t = self.start_clock("Compose and find shortest path")
        best_path = self.viterbi_search_batch([candidates], [distances])[0]
        self.stop_clock(t)
        return best_path

    def viterbi_search_batch(self, candidates_list, distances_list, return_costs=False, greedy=False):
        self._push_weights()
        flags = engine.VITERBI_BEAM1 if greedy else 0
        paths, pcost, tcost, jcost = self.db.join_viterbi_batch(
            [np.asarray(c) for c in candidates_list], [np.asarray(d) for d in distances_list], flags)
        if return_costs:
            return paths, pcost, tcost, jcost
        return paths

    def synthesise_paths_acoustic_batch(self, unit_features_list, return_costs=False):
        """preselect_units_acoustic -> viterbi_search of synth_utt (synth_halfphone.py:1611-1625) for a batch of
        utterances in ONE engine call: the candidate lists never leave the device, only the paths come back."""
        self._push_weights()
        feats = [np.asarray(u, dtype=np.float64) for u in unit_features_list]
        lens = np.array([u.shape[0] for u in feats], dtype=np.int64)
        paths, pcost, tcost, jcost = self.db.acoustic_viterbi_batch_cat(np.concatenate(feats, axis=0), lens,
                                                                        self.config["n_candidates"])
        if return_costs:
            return paths, pcost, tcost, jcost
        return paths

    # ---- cost report (synth_halfphone.py:1964-1981)
    def get_scores_per_stream(self, unit_features, best_path):
        self._push_weights()
        tw = [self.datadims_target[s] for s in self.stream_list_target]
        jw = [self.datadims_join[s] for s in self.stream_list_join]
        return self.db.greedy_path_scores(np.asarray(unit_features, dtype=np.float64), best_path, tw, jw)

    def get_target_scores_per_stream(self, target_features, best_path):
        return self.get_scores_per_stream(target_features, best_path)[0]

    def get_join_scores_per_stream(self, best_path, target_features=None):
        if target_features is None:
            target_features = np.zeros((len(best_path) * self._multiepoch, self.db.Dt))
        return self.get_scores_per_stream(target_features, best_path)[1]
