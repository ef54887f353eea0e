"""CPU oracle for snickery's unit-selection search path -- TEST INFRASTRUCTURE ONLY.

This module restates, in Python 3 + numpy/scipy, the algorithm of the reference's
hot path (weighting -> k-NN candidate query -> join costs -> greedy / Viterbi
search).  It exists so the CUDA path can be checked; nothing in the product
(`snickery_b200/`) may import it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may execute it.

PARITY PINNING.  The reference (Python 2.7, needs h5py / pywrapfst / magphase)
cannot be imported in this image and ships NO tests, golden vectors or
known-answer fixtures for this path.  The oracle is pinned by
  (1) OUTPUTS OF THE REFERENCE'S OWN CODE: oracle/ref_exec.py executes the reference's
      functions for this path (mechanical Python 2 -> 3 transform of the files where they
      lie) on seeded inputs; the results are committed as tests/golden/reference_exec.npz and
      tests/test_reference_exec.py requires this module to reproduce every one of them;
  (2) the reference's third-party engines where they exist here: `scipy.spatial.cKDTree`
      with the reference's constructor/query kwargs (script/synth_simple.py:229,490;
      script/synth_halfphone.py:379,1364) and `sklearn.neighbors.KDTree`;
  (3) the reference's only in-code known answers: the natural-path identity assertion
      (script/synth_simple.py:909-928, script/synth_halfphone.py:1455-1474) and zero join
      cost between adjacent units (script/synth_simple.py:250-251);
  (4) exhaustive path enumeration on small lattices for the Viterbi restatement.
OpenFst 1.5.4 / pywrapfst (pinned in README_FULL.md:45-53) is absent: the reference's calls
into it land in oracle/minifst.py, a restatement of the published algorithms.  What stays
"parity unpinned" is OpenFst's own tie order between equal-cost paths; DESIGN.md says so.

Every function cites the reference file:line it follows (paths under
/root/reference/script/).
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.spatial

VERY_BIG_WEIGHT_VALUE = 1000000000000000.0  # const.py:3
TARGET_REP_WIDTHS = {"onepoint": 1, "twopoint": 2, "threepoint": 3, "epoch": 1, "sample": 1}  # const.py:17
APPLY_JCW_ON_TOP = True  # synth_simple.py:44


# --------------------------------------------------------------------------- helpers
def weight(speech, weight_vec):
    """speech_manip.py:209-213 -- f32 array times f64 row vector -> f64."""
    weight_vec = np.array(weight_vec, dtype=np.float64).reshape((1, -1))
    return speech * weight_vec


def halfphone_stats(speech, labels, representation_type="twopoint"):
    """train_halfphone.py:959-1070 -- (names, features, timings) of the half-phones of one utterance.  labels: list of
    ((start, end), [ll, l, c, r, rr, state]), five states '2'..'6' per phone; states 2-3 are the left half-phone, 4-6
    the right one; a unit is described by its first frame (start of state 2 / 4), its middle frame (end of state 2 / 5)
    and its last frame (end of state 3 / 6), ends clipped to the utterance."""
    m = speech.shape[0]
    assert len(labels) % 5 == 0
    names, starts, middles, ends = [], [], [], []
    for (s, e), lab in labels:
        e = min(e, m - 1)
        quin, state = list(lab[:5]), lab[-1]
        if state == "2":
            quin[2] += "_L"
            names.append("/".join(quin)); starts.append(s); middles.append(e)
        elif state == "3":
            ends.append(e)
        elif state == "4":
            quin[2] += "_R"
            names.append("/".join(quin)); starts.append(s)
        elif state == "5":
            middles.append(e)
        elif state == "6":
            ends.append(e)
        else:
            raise SystemExit("bad state number")
    if representation_type == "onepoint":
        feats = speech[middles, :]
    elif representation_type == "twopoint":
        feats = np.hstack([speech[starts, :], speech[ends, :]])
    else:
        feats = np.hstack([speech[starts, :], speech[middles, :], speech[ends, :]])
    return np.array(names), feats, list(zip(starts, ends))


SPECIAL_UV_VALUE = -1000.0   # const.py:12
UV_SCALING_FACTOR = 20.0     # const.py:14


def standardise(speech, mean_vec, std_vec):
    """data_manipulation.py:162-186 -- (speech - mean) / std; entries that held the unvoiced marker before
    standardisation become std * -1.0 * uv_scaling_factor.  Weighting is left to weight().
    The arithmetic type is numpy's: the voice file stores mean / std as float32 (train_simple.py:94-97), so with
    float32 speech everything here stays float32; float64 statistics promote it to float64."""
    speech = np.asarray(speech)
    uv_positions = (speech == SPECIAL_UV_VALUE)
    mean_vec = np.asarray(mean_vec).reshape((1, -1))
    std_vec = np.asarray(std_vec).reshape((1, -1))
    speech = (speech - mean_vec) / std_vec
    uv_values = std_vec * -1.0 * UV_SCALING_FACTOR
    for column in range(speech.shape[1]):
        speech[:, column][uv_positions[:, column]] = uv_values[0, column]
    return speech


def segment_axis0(a, length, overlap):
    """segmentaxis.py:40-111 restricted to axis=0, end='cut'.

    Returns [n, length, ...] windows with hop length-overlap; a remainder that
    does not fill a window is cut; fewer than `length` rows is a ValueError
    (segmentaxis.py:94-96)."""
    if overlap >= length:
        raise ValueError("frames cannot overlap by more than 100%")
    if overlap < 0 or length <= 0:
        raise ValueError("overlap must be nonnegative and length must be positive")
    hop = length - overlap
    l = a.shape[0]
    if l < length:
        raise ValueError("Not enough data points to segment array in 'cut' mode")
    n = 1 + (l - length) // hop
    idx = (np.arange(n) * hop)[:, None] + np.arange(length)[None, :]
    return a[idx]


def per_coeff_weights(stream_weights, stream_list, datadims, nrepetitions=1, extra=None):
    """synth_simple.py:236-243 / 259-267; synth_halfphone.py:713-726."""
    assert len(stream_weights) == len(stream_list), (stream_weights, stream_list)
    vec = []
    for i, stream in enumerate(stream_list):
        vec.extend([float(stream_weights[i])] * datadims[stream])
    vec = vec * nrepetitions
    if extra is not None:
        vec.append(float(extra))
    return np.array(vec, dtype=np.float64)


# --------------------------------------------------------------------------- synthesiser
class OracleSynthesiser:
    """Restates the search-relevant state of synth_simple.Synthesiser /
    synth_halfphone.Synthesiser.  Arrays are given instead of read from HDF5."""

    def __init__(self, config, F, Jc):
        self.config = dict(config)
        self.train_unit_features_unweighted = np.asarray(F, dtype=np.float32)
        self.join_contexts_unweighted = np.asarray(Jc, dtype=np.float32)
        self.stream_list_target = config["stream_list_target"]
        self.stream_list_join = config["stream_list_join"]
        self.datadims_target = config["datadims_target"]
        self.datadims_join = config["datadims_join"]
        self.target_representation = config.get("target_representation", "epoch")
        jcw = config["join_cost_weight"]
        # synth_simple.py:128-133 / synth_halfphone.py:260-270
        if APPLY_JCW_ON_TOP:
            self.set_target_weights(np.array(config["target_stream_weights"]) * (1.0 - jcw))
            self.set_join_weights(np.array(config["join_stream_weights"]) * jcw)
        else:
            self.set_target_weights(config["target_stream_weights"])
            self.set_join_weights(config["join_stream_weights"])

    # ---- W1 / W2
    def set_join_weights(self, weights):
        """synth_simple.py:234-255; synth_halfphone.py:682-707 doubles the vector for epoch voices written
        by train_halfphone.py (two-frame join windows, :693-695)."""
        w = per_coeff_weights(weights, self.stream_list_join, self.datadims_join)
        if self.config.get("halfphone_epoch_join_layout", False):
            w = np.concatenate([w, w])
        jw = weight(self.join_contexts_unweighted, w)
        self.join_weight_vector = w
        self.unit_end_data = jw[1:, :]
        self.unit_start_data = jw[:-1, :]

    def set_target_weights(self, weights):
        """synth_simple.py:257-274; synth_halfphone.py:713-737."""
        extra = None
        if self.config.get("add_duration_as_target", False):
            extra = self.config.get("duration_target_weight", 0.0)
        w = per_coeff_weights(weights, self.stream_list_target, self.datadims_target,
                              TARGET_REP_WIDTHS[self.target_representation], extra)
        self.train_unit_features = weight(self.train_unit_features_unweighted, w)
        self.target_weight_vector = w

    # ---- G1
    def combined_rep(self):
        """synth_simple.py:190-224: [prev_join_rep || windowed target features].
        With halfphone_epoch_join_layout the join contexts hold two frames per row and
        prev / current are the two halves of unit_start_data (synth_halfphone.py:552-553)."""
        if self.config.get("halfphone_epoch_join_layout", False):
            n = self.unit_start_data.shape[1]
            self.prev_join_rep = self.unit_start_data[:, :n // 2]
            self.current_join_rep = self.unit_start_data[:, n // 2:]
        else:
            self.prev_join_rep = self.unit_start_data
            self.current_join_rep = self.unit_end_data
        feats = self.train_unit_features
        m_ep = self.config.get("multiepoch", 1)
        if m_ep > 1:
            overlap = m_ep - 1
            m, n = feats.shape
            feats = segment_axis0(feats, m_ep, overlap).reshape(m - overlap, n * m_ep)
            self.current_join_rep = self.current_join_rep[overlap:, :]
            self.prev_join_rep = self.prev_join_rep[:-overlap, :]
        self.windowed_unit_features = feats
        return np.hstack([self.prev_join_rep, feats])

    def get_tree_for_greedy_search(self):
        """synth_simple.py:226-230."""
        combined = self.combined_rep()
        self.joint_tree = scipy.spatial.cKDTree(combined, leafsize=100, balanced_tree=False)
        return self.joint_tree

    # ---- G2
    def window_targets(self, unit_features):
        """synth_simple.py:473-478."""
        m_ep = self.config.get("multiepoch", 1)
        if m_ep > 1:
            m, n = unit_features.shape
            unit_features = segment_axis0(unit_features, m_ep, 0).reshape(m // m_ep, n * m_ep)
        return unit_features

    def greedy_joint_search(self, unit_features, start_state=-1, engine="tree", return_dists=False):
        """synth_simple.py:458-503.  engine='tree' uses the cKDTree exactly as the
        reference does (eps forced to 0 = exact search); engine='brute' is a
        float64 brute-force argmin with lowest-index tie rule."""
        n = self.current_join_rep.shape[1]
        if start_state < 0:
            prev = np.zeros((n,))
        else:
            prev = self.prev_join_rep[start_state, :]
        unit_features = self.window_targets(np.asarray(unit_features, dtype=np.float64))
        if engine == "brute":
            combined = np.hstack([self.prev_join_rep, self.windowed_unit_features])
        path, dists_out = [], []
        for target_vector in unit_features:
            both = np.concatenate([prev, target_vector]).reshape((1, -1))
            if engine == "tree":
                dists, indexes = self.joint_tree.query(both, k=1, eps=0.0)
                d, ix = float(dists.flatten()[0]), int(indexes.flatten()[0])
            else:
                d2 = ((combined - both) ** 2).sum(axis=1)
                ix = int(np.argmin(d2))
                d = float(np.sqrt(d2[ix]))
            path.append(ix)
            dists_out.append(d)
            prev = self.current_join_rep[ix, :]
        if return_dists:
            return path, np.array(dists_out)
        return path

    def greedy_step_distances(self, prev_vector, target_vector):
        """All joint distances for one greedy step (float64), for tie auditing."""
        combined = np.hstack([self.prev_join_rep, self.windowed_unit_features])
        both = np.concatenate([prev_vector, target_vector])[None, :]
        return np.sqrt(((combined - both) ** 2).sum(axis=1))

    # ---- K1
    def build_acoustic_tree(self):
        """synth_halfphone.py:345-382 (:379)."""
        self.tree = scipy.spatial.cKDTree(self.train_unit_features, leafsize=100,
                                          compact_nodes=False, balanced_tree=False)
        return self.tree

    def preselect_units_acoustic(self, unit_features):
        """synth_halfphone.py:1359-1366."""
        distances, candidates = self.tree.query(unit_features, k=self.config["n_candidates"])
        return candidates, distances

    # ---- W3
    def get_selection_vector(self, stream_list, stream_dims, truncation_values):
        """synth_simple.py:968-980."""
        assert len(truncation_values) == len(stream_list)
        sel, start = [], 0
        for stream, trunc in zip(stream_list, truncation_values):
            dim = stream_dims[stream]
            if trunc == -1:
                trunc = dim
            assert trunc <= dim
            sel.extend(range(start, start + trunc))
            start += dim
        return sel

    def truncate_join_streams(self, truncation_values):
        """synth_simple.py:982-985."""
        sel = self.get_selection_vector(self.stream_list_join, self.datadims_join, truncation_values)
        self.unit_end_data = self.unit_end_data[:, sel]
        self.unit_start_data = self.unit_start_data[:, sel]

    def truncate_target_streams(self, truncation_values):
        """synth_simple.py:987-992."""
        sel = self.get_selection_vector(self.stream_list_target, self.datadims_target, truncation_values)
        self.train_unit_features = self.train_unit_features[:, sel]
        self.target_truncation_vector = sel

    # ---- K2
    def build_phonetrees(self, train_unit_names):
        """synth_halfphone.py:385-402."""
        monophones = np.array([q.split("/")[2] for q in train_unit_names])
        self.phonetrees, self.phonetrees_index_converters = {}, {}
        n = self.train_unit_features.shape[0]
        for phone in dict.fromkeys(monophones.tolist()):
            sel = monophones == phone
            self.phonetrees[phone] = scipy.spatial.cKDTree(self.train_unit_features[sel, :], leafsize=10,
                                                           compact_nodes=False, balanced_tree=False)
            self.phonetrees_index_converters[phone] = np.arange(n)[sel]

    def preselect_units_monophone_then_acoustic(self, unit_features, unit_names):
        """synth_halfphone.py:1369-1396 (phones with fewer than K units keep the -1 / 1e15 padding)."""
        K = self.config["n_candidates"]
        m = unit_features.shape[0]
        candidates = np.ones((m, K), dtype=int) * -1
        distances = np.ones((m, K)) * VERY_BIG_WEIGHT_VALUE
        monophones = [q.split("/")[2] for q in unit_names]
        for i, phone in enumerate(monophones):
            conv = self.phonetrees_index_converters[phone]
            kk = min(K, conv.size)
            d, c = self.phonetrees[phone].query(unit_features[i, :], k=kk)
            d, c = np.atleast_1d(d), np.atleast_1d(c)
            candidates[i, :kk] = conv[c]
            distances[i, :kk] = d
        return candidates, distances

    # ---- K3 (candidate half)
    def preselect_units_quinphone(self, unit_features, unit_names, unit_index):
        """synth_halfphone.py:1305-1354 with label_manip.py:16-32."""
        K = self.config["n_candidates"]
        candidates = []
        for quinphone in unit_names:
            q = quinphone.split("/")
            mono, tri = q[2], "/".join(q[1:4])
            di = "/".join(q[1:3]) if mono.endswith("_L") else "/".join(q[2:4])
            current = []
            for form in [quinphone, tri, di, mono]:
                for unit in unit_index.get(form, []):
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
        return candidates, self.candidate_distances(candidates, unit_features)

    # ---- K3 (distance half)
    def candidate_distances(self, candidates, unit_features):
        """synth_halfphone.py:1346-1351 -- rows for -1 index the LAST unit (numpy
        negative indexing); the T lattice skips them later."""
        distances = []
        for i, row in enumerate(candidates):
            cf = self.train_unit_features[row]
            tf = unit_features[i].reshape((1, -1))
            distances.append(np.sqrt(np.sum((cf - tf) ** 2, axis=1)))
        return np.array(distances)

    # ---- J1
    def get_natural_distance_vectorised(self, first, second, order=1):
        """synth_halfphone.py:2942-2951."""
        sq = (self.unit_end_data[first, :] - self.unit_start_data[second, :]) ** 2
        return (1.0 / order) * np.sqrt(np.sum(sq, axis=1))

    # ---- J2
    def join_cost_cache(self, ind):
        """synth_halfphone.py:3206-3301: the (first, second) -> cost dict of
        make_on_the_fly_join_lattice_BLOCK_DIRECT (order=1)."""
        data_frames = self.unit_end_data.shape[0]
        mini, maxi = 1, data_frames - 1  # :3238-3240
        frames, _ = np.shape(ind)
        first_list, second_list = [], []
        for i in range(frames - 1):
            for first in ind[i, :]:
                if first < mini or first >= maxi:
                    continue
                for second in ind[i + 1, :]:
                    if second < mini or second >= maxi:
                        continue
                    if first == -1 or second == -1:
                        continue
                    first_list.append(int(first))
                    second_list.append(int(second))
        if not first_list:
            return {}
        dists = self.get_natural_distance_vectorised(first_list, second_list, order=1)
        return dict(((l, r), w) for l, r, w in zip(first_list, second_list, dists))

    # ---- V1-V4
    def viterbi_search(self, candidates, distances, arithmetic="f64", return_cost=False):
        """Net semantics of make_target_sausage_lattice (fst_functions_wrapped.py:28-58),
        cost_cache_to_compiled_fst (:172-217), compose (:368) and tropical 1-best
        shortestpath (:387-408), as an explicit min-plus DP.

        A path u_0..u_{T-1} exists iff u_t is a non -1 candidate at t and every
        consecutive pair is in the join cost cache.  The composed arc at step t
        carries D[t,u_t] (+) join(u_t,u_{t+1}); the last arc carries D[T-1,u] + 0.

        arithmetic='f64'      : float64 throughout (numpy's natural arithmetic)
        arithmetic='openfst32': every weight goes through Python-2 '%s' text
                                (12 significant digits, fst_functions_wrapped.py:47,201)
                                into TropicalWeight<float>; Times is a float32 add
                                applied arc-by-arc in forward order.
        Ties: lowest (t, candidate column) predecessor first -- OpenFst's own
        tie order is unspecified (queue order)."""
        candidates = np.asarray(candidates)
        distances = np.asarray(distances, dtype=np.float64)
        T, K = candidates.shape
        cache = self.join_cost_cache(candidates)
        if T < 2 or not cache:
            return ([], np.inf) if return_cost else []
        if arithmetic == "openfst32":
            conv = lambda x: np.float32(float("%.12g" % x))
            add = lambda a, b: np.float32(np.float32(a) + np.float32(b))
            inf = np.float32(np.inf)
        else:
            conv = float
            add = lambda a, b: a + b
            inf = np.inf
        states = set()
        for (a, b) in cache:
            states.add(a)
            states.add(b)
        # d[j] = best cost of reaching J-state of candidate j at time t (arcs 0..t-1 consumed)
        d = [inf] * K
        for j in range(K):
            u = int(candidates[0, j])
            if u != -1 and u in states:
                d[j] = conv(0.0)
        bp = np.full((T, K), -1, dtype=np.int64)
        for t in range(T - 1):
            nd = [inf] * K
            for jb in range(K):
                b = int(candidates[t + 1, jb])
                if b == -1:
                    continue
                best, arg = inf, -1
                for ja in range(K):
                    if d[ja] == inf:
                        continue
                    a = int(candidates[t, ja])
                    w = cache.get((a, b))
                    if w is None:
                        continue
                    arc = add(conv(distances[t, ja]), conv(w))
                    c = add(d[ja], arc)
                    if c < best:
                        best, arg = c, ja
                nd[jb] = best
                bp[t + 1, jb] = arg
            d = nd
        best, arg = inf, -1
        for j in range(K):
            if d[j] == inf:
                continue
            c = add(d[j], add(conv(distances[T - 1, j]), conv(0.0)))
            if c < best:
                best, arg = c, j
        if arg < 0:
            return ([], np.inf) if return_cost else []
        cols = [arg]
        for t in range(T - 1, 0, -1):
            cols.append(int(bp[t, cols[-1]]))
        cols.reverse()
        path = [int(candidates[t, j]) for t, j in enumerate(cols)]
        if return_cost:
            return path, float(best)
        return path

    def viterbi_exhaustive(self, candidates, distances):
        """Brute-force enumeration of every admissible path (float64); only for
        tiny lattices.  Independent check of viterbi_search."""
        candidates = np.asarray(candidates)
        T, K = candidates.shape
        cache = self.join_cost_cache(candidates)
        best, best_path = np.inf, []
        for cols in itertools.product(range(K), repeat=T):
            units = [int(candidates[t, j]) for t, j in enumerate(cols)]
            if -1 in units:
                continue
            cost, ok = 0.0, True
            for t in range(T - 1):
                w = cache.get((units[t], units[t + 1]))
                if w is None:
                    ok = False
                    break
                cost += distances[t, cols[t]] + w
            if not ok:
                continue
            cost += distances[T - 1, cols[T - 1]]
            if cost < best:
                best, best_path = cost, units
        return best_path, best

    def path_costs(self, candidates, distances, path):
        """Float64 target / join / total cost of a given unit path through a lattice
        (first matching candidate column per step)."""
        tc, jc = 0.0, 0.0
        for t, u in enumerate(path):
            cols = np.flatnonzero(np.asarray(candidates[t]) == u)
            assert len(cols) > 0, "unit %d is not a candidate at step %d" % (u, t)
            tc += float(np.min(np.asarray(distances)[t, cols]))
        if len(path) > 1:
            p = np.asarray(path)
            jc = float(self.get_natural_distance_vectorised(p[:-1], p[1:], order=1).sum())
        return tc, jc, tc + jc

    # ---- C1
    def aggregate_squared_errors_by_stream(self, squared_errors, cost_type):
        """synth_halfphone.py:2977-3008 (no sqrt)."""
        if cost_type == "target":
            streams, widths = self.stream_list_target, self.datadims_target
        else:
            streams, widths = self.stream_list_join, self.datadims_join
        m, _ = squared_errors.shape
        out = np.ones((m, len(streams))) * -1.0
        start = 0
        for i, s in enumerate(streams):
            out[:, i] = np.sum(squared_errors[:, start:start + widths[s]], axis=1)
            start += widths[s]
        return out

    def get_target_scores_per_stream(self, target_features, best_path):
        """synth_halfphone.py:1964-1969 (epoch greedy: features are the windowed ones)."""
        feats = getattr(self, "windowed_unit_features", self.train_unit_features)
        chosen = feats[best_path]
        return self.aggregate_squared_errors_by_stream((chosen - target_features) ** 2, "target")

    def get_join_scores_per_stream(self, best_path):
        """synth_halfphone.py:1971-1981."""
        bp = np.array(best_path)
        if self.config.get("greedy_search", False):
            sq = (self.prev_join_rep[bp[1:], :] - self.current_join_rep[bp[:-1], :]) ** 2
        else:
            sq = (self.unit_end_data[bp[:-1], :] - self.unit_start_data[bp[1:], :]) ** 2
        return self.aggregate_squared_errors_by_stream(sq, "join")


# --------------------------------------------------------------------------- plain k-NN oracle
def brute_force_knn(data, queries, k):
    """float64 exact k-NN (Euclidean, ascending, lowest index on ties)."""
    data = np.asarray(data, dtype=np.float64)
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float64))
    dd = np.empty((queries.shape[0], k))
    ii = np.empty((queries.shape[0], k), dtype=np.int64)
    for q in range(queries.shape[0]):
        d2 = ((data - queries[q][None, :]) ** 2).sum(axis=1)
        order = np.lexsort((np.arange(len(d2)), d2))[:k]
        ii[q] = order
        dd[q] = np.sqrt(d2[order])
    return dd, ii


def viterbi_search_numpy(synth, candidates, distances, return_cost=False):
    """Vectorised float64 version of OracleSynthesiser.viterbi_search (same
    semantics, same lowest-column tie rule): the K x K join tiles are computed
    with the numpy gather of synth_halfphone.py:2942-2951 and masked with the
    admissibility rules of :3238-3268.  Used for larger parity cases and as the
    CPU baseline's search stage."""
    cand = np.asarray(candidates).astype(np.int64)
    D = np.asarray(distances, dtype=np.float64)
    T, K = cand.shape
    end, start = synth.unit_end_data, synth.unit_start_data
    n = end.shape[0]
    ok = (cand >= 1) & (cand < n - 1)            # covers -1 as well
    if T < 2:
        return ([], np.inf) if return_cost else []
    safe = np.where(ok, cand, 1)
    # a unit is a J state iff it occurs in some admissible pair
    has_next = np.zeros((T, K), bool)
    has_prev = np.zeros((T, K), bool)
    has_next[:-1] = ok[:-1] & ok[1:].any(axis=1, keepdims=True)
    has_prev[1:] = ok[1:] & ok[:-1].any(axis=1, keepdims=True)
    in_pairs = has_next | has_prev
    state_units = set(cand[in_pairs].tolist())
    d = np.where([(int(u) in state_units) and u != -1 for u in cand[0]], 0.0, np.inf)
    bp = np.full((T, K), -1, dtype=np.int64)
    for t in range(T - 1):
        e = end[safe[t]]                          # [K, Dj]
        s = start[safe[t + 1]]
        J = np.sqrt(((e[:, None, :] - s[None, :, :]) ** 2).sum(axis=2))
        J[~ok[t], :] = np.inf
        J[:, ~ok[t + 1]] = np.inf
        tot = d[:, None] + (D[t][:, None] + J)
        arg = np.argmin(tot, axis=0)
        nd = tot[arg, np.arange(K)]
        bp[t + 1] = np.where(np.isfinite(nd), arg, -1)
        d = nd
    fin = d + (D[T - 1] + 0.0)
    arg = int(np.argmin(fin))
    if not np.isfinite(fin[arg]):
        return ([], np.inf) if return_cost else []
    cols = [arg]
    for t in range(T - 1, 0, -1):
        cols.append(int(bp[t, cols[-1]]))
    cols.reverse()
    path = [int(cand[t, j]) for t, j in enumerate(cols)]
    return (path, float(fin[arg])) if return_cost else path


# --------------------------------------------------------------------------- row N2: MagPhase epoch concatenation
def zero_pad_matrix(a, start_pad, end_pad):
    """matrix_operations.py:3-13."""
    if start_pad > 0:
        a = np.vstack([np.zeros((start_pad, a.shape[1])), a])
    if end_pad > 0:
        a = np.vstack([a, np.zeros((end_pad, a.shape[1]))])
    return a


def taper_matrix(a, taper_length):
    """matrix_operations.py:16-30 -- multiplies IN PLACE (a float32 slice stays float32)."""
    m, n = a.shape
    assert taper_length * 2 <= m, "taper_length (%s) too long for (padded) unit length (%s)" % (taper_length, m)
    in_taper = np.hanning(((taper_length + 1) * 2) + 1)[1:taper_length + 1].reshape(-1, 1)
    out_taper = np.flipud(in_taper).reshape(-1, 1)
    a[:taper_length, :] *= in_taper
    a[-taper_length:, :] *= out_taper
    return a


def lin_interp_f0(fz):
    """speech_manip.py:222-244."""
    import scipy.interpolate
    y = fz.flatten()
    voiced_ix = np.where(y > 0.0)[0]
    voicing_flag = np.zeros(y.shape)
    voicing_flag[voiced_ix] = 1.0
    if voiced_ix.shape[0] == 0:
        v_interpolated = fz
    else:
        interpolator = scipy.interpolate.interp1d(voiced_ix, y[voiced_ix], kind="linear", axis=0,
                                                  bounds_error=False, fill_value="extrapolate")
        v_interpolated = interpolator(np.arange(y.shape[0]))
    return (v_interpolated.reshape((-1, 1)), voicing_flag.reshape((-1, 1)))


class OracleMagPhaseStore:
    """The per-sentence full-band files as retrieve_magphase_frag reads them (synth_simple.py:538-563):
    sentences[name] = (mag, real, imag, f0) float32 arrays; unit u lives in train_filenames[u] at
    unit_index_within_sentence[u].  Files are re-read (copied) on every retrieval like get_speech does."""

    def __init__(self, sentences, train_filenames, unit_index_within_sentence, multiepoch=1):
        self.sentences = sentences
        self.train_filenames = train_filenames
        self.unit_index_within_sentence = unit_index_within_sentence
        self.multiepoch = multiepoch

    def retrieve_magphase_frag(self, index, extra_frames=0):
        """synth_simple.py:538-652."""
        mag_full, real_full, imag_full, f0_full = (a.copy() for a in self.sentences[self.train_filenames[index]])
        f0_interp, vuv = lin_interp_f0(f0_full)
        start_index = self.unit_index_within_sentence[index]
        end_index = start_index + self.multiepoch
        start_pad = end_pad = 0
        if extra_frames > 0:
            new_start_index = start_index - extra_frames
            new_end_index = end_index + extra_frames
            nframes, _ = mag_full.shape
            if new_start_index < 0:
                start_pad = new_start_index * -1
            if new_end_index > nframes:
                end_pad = new_end_index - nframes
            start_index = 0 if start_pad > 0 else new_start_index
            end_index = nframes if end_pad > 0 else new_end_index
        frags = [a[start_index:end_index, :] for a in (mag_full, real_full, imag_full, f0_interp, vuv)]
        frags = [zero_pad_matrix(a, start_pad, end_pad) for a in frags]
        assert frags[0].shape[0] == self.multiepoch + extra_frames * 2
        if extra_frames > 0:
            frags = [taper_matrix(a, extra_frames * 2) for a in frags]
        return tuple(frags)

    def concatenate(self, path, overlap=0, fzero=np.zeros(0)):
        """synth_simple.py:677-729 up to (not including) magphase.synthesis_from_lossless."""
        assert overlap % 2 == 0, "frame overlap should be even number"
        m = self.multiepoch
        width = next(iter(self.sentences.values()))[0].shape[1]
        nframes = len(path) * m + overlap
        mag, real, imag = np.zeros((nframes, width)), np.zeros((nframes, width)), np.zeros((nframes, width))
        fz, vuv = np.zeros((nframes, 1)), np.zeros((nframes, 1))
        write_start = 0
        for ix in path:
            write_end = write_start + m + overlap
            mag_frag, real_frag, imag_frag, fz_frag, vuv_frag = self.retrieve_magphase_frag(ix, extra_frames=overlap // 2)
            mag[write_start:write_end, :] += mag_frag
            real[write_start:write_end, :] += real_frag
            imag[write_start:write_end, :] += imag_frag
            fz[write_start:write_end, :] += fz_frag
            vuv[write_start:write_end, :] += vuv_frag
            write_start += m
        if overlap > 0:
            taper = overlap // 2
            mag, real, imag, fz, vuv = (a[taper:-taper, :] for a in (mag, real, imag, fz, vuv))
        if fzero.size > 0:
            fz = fzero
        else:
            unvoiced = np.where(vuv < 0.5)[0]
            fz[unvoiced, :] = 0.0
        return mag, real, imag, fz
