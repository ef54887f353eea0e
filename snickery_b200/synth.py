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


class Synthesiser:
    def __init__(self, config, train_unit_features_unweighted, join_contexts_unweighted, device=0, verbose=False):
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
        jcw = self.config["join_cost_weight"]
        if APPLY_JCW_ON_TOP:   # synth_simple.py:128-133
            self.set_target_weights(np.array(self.config["target_stream_weights"]) * (1.0 - jcw))
            self.set_join_weights(np.array(self.config["join_stream_weights"]) * jcw)
        else:
            self.set_target_weights(self.config["target_stream_weights"])
            self.set_join_weights(self.config["join_stream_weights"])
        self._push_weights()   # the reference weights its arrays in the constructor (synth_simple.py:128-133)
        if greedy_epoch:
            self.get_tree_for_greedy_search()
        elif self.config.get("preselection_method", "quinphone") == "acoustic":
            self.tree = GpuKDTree(None, _db=self.db, _space=engine.SPACE_TARGET)

    # ---- clocks, as the reference prints them (synth_simple.py:760-768)
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
            self.db.set_weights(self._wt, self._wj)
            self._dirty = False
            self.stop_clock(t)

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
        feats = [np.asarray(u, dtype=np.float64) for u in unit_features_list]
        return self.db.greedy_batch(feats, start_states, return_dists=return_dists)

    # ---- preselection (synth_halfphone.py:1359-1366, 1346-1351)
    def preselect_units_acoustic(self, unit_features):
        self._push_weights()
        t = self.start_clock("Acoustic select units ")
        distances, candidates = self.tree.query(unit_features, k=self.config["n_candidates"])
        self.stop_clock(t)
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
