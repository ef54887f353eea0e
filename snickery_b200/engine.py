"""ctypes binding of the C ABI in include/snk_b200.h.

This is the only place the Python host touches the CUDA engine.  There is no CPU
fallback: if the shared library is missing or no B200 is visible, construction of
a UnitDatabase raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsnk_b200.so")

LAYOUT_SIMPLE = 0
LAYOUT_HALFPHONE_EPOCH = 1
SPACE_TARGET = 0
SPACE_JOINT = 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
VITERBI_BEAM1 = 1
PROF_KNN, PROF_JOIN, PROF_VITERBI, PROF_ALLGATHER, PROF_MERGE, PROF_JOIN_VITERBI, PROF_RERANK = 0, 1, 2, 3, 4, 5, 6

_lib = None


class EngineError(RuntimeError):
    pass


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


def load_library(path=None):
    """Loads libsnk_b200.so (building is the job of `python -m snickery_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError("CUDA engine not built: %s is missing (run `python -m snickery_b200.build`). "
                          "There is no CPU fallback." % path)
    lib = C.CDLL(path)
    i64, i32, dbl, flt, vp = C.c_int64, C.c_int, C.c_double, C.c_float, C.c_void_p
    P = C.POINTER
    lib.snk_last_error.restype = C.c_char_p
    lib.snk_version.restype = i32
    lib.snk_device_count.restype = i32
    sigs = {
        "snk_db_create": [P(vp), i32, i64, i32, i32, i32, P(flt), P(flt), C.c_uint],
        "snk_db_destroy": [vp],
        "snk_db_info": [vp, P(i64), P(i64), P(i32), P(i32), P(i32), P(i32)],
        "snk_db_set_weights": [vp, P(dbl), P(dbl)],
        "snk_db_set_engine": [vp, i32],
        "snk_db_counters": [vp, P(i64), i32],
        "snk_db_profile_enable": [vp, i32],
        "snk_db_profile_read": [vp, i32, P(dbl), P(i64), P(dbl), i32],
        "snk_knn": [vp, i32, P(dbl), i64, i32, P(dbl), P(i64)],
        "snk_knn_dev": [vp, i32, vp, i64, i32, vp, vp, i64, vp],
        "snk_topk_merge_dev": [i32, vp, vp, i32, i64, i32, vp, vp, vp],
        "snk_knn_finish": [vp],
        "snk_debug_tc_keys": [vp, i32, P(dbl), i64, i64, i64, P(flt), P(flt), P(flt), P(flt)],
        "snk_debug_greedy_one_keys": [vp, P(dbl), i64, P(flt), P(flt), P(flt), P(flt)],
        "snk_debug_greedy_one_times": [vp, vp, i32],
        "snk_debug_greedy_one_cta_times": [vp, vp, i32],
        "snk_greedy_batch_finish": [vp],
        "snk_comm_unique_id": [vp, i32],
        "snk_comm_init": [vp, vp, i32, i32],
        "snk_comm_peer_exchange": [vp],
        "snk_comm_info": [vp, P(i32), P(i32), P(i32)],
        "snk_knn_sharded_dev": [vp, i32, vp, i64, i32, vp, vp, i64, vp],
        "snk_knn_sharded_finish": [vp],
        "snk_greedy_batch": [vp, P(dbl), P(i64), i32, P(i64), P(i64), P(dbl)],
        "snk_greedy_batch_dev": [vp, vp, P(i64), i32, P(i64), vp, vp, vp],
        "snk_greedy_sharded_batch_dev": [vp, vp, P(i64), i32, P(i64), vp, i64, i64, vp, vp, vp],
        "snk_db_set_standardisation": [vp, P(dbl), P(dbl), dbl, dbl, C.c_uint],
        "snk_prepare_targets": [vp, P(flt), i64, P(dbl)],
        "snk_halfphone_targets": [vp, P(flt), i64, i32, P(i64), i64, i32, P(dbl), P(dbl)],
        "snk_halfphone_targets_dev": [vp, vp, i64, i32, vp, i64, i32, vp, vp, vp],
        "snk_greedy_batch_unnorm": [vp, P(flt), P(i64), i32, P(i64), P(i64), P(dbl)],
        "snk_greedy_batch_unnorm_dev": [vp, vp, P(i64), i32, P(i64), vp, vp, vp],
        "snk_candidate_distances": [vp, P(i64), P(dbl), i64, i32, P(dbl)],
        "snk_join_tiles": [vp, P(i64), P(i64), i32, i32, P(flt)],
        "snk_join_stats": [vp, P(i64)],
        "snk_join_viterbi_batch": [vp, P(i64), P(dbl), P(i64), i32, i32, C.c_uint, P(i64), P(i64), P(dbl), P(dbl), P(dbl)],
        "snk_join_viterbi_batch_dev": [vp, vp, vp, P(i64), i32, i32, C.c_uint, vp, vp, vp, vp, vp, vp],
        "snk_acoustic_viterbi_batch": [vp, P(dbl), P(i64), i32, i32, C.c_uint, P(i64), P(i64), P(dbl), P(dbl), P(dbl)],
        "snk_acoustic_viterbi_batch_dev": [vp, vp, P(i64), i32, i32, C.c_uint, vp, vp, vp, vp, vp, vp],
        "snk_acoustic_viterbi_finish": [vp],
        "snk_greedy_path_scores": [vp, P(dbl), i64, P(i64), i64, P(i32), i32, P(i32), i32, P(dbl), P(dbl)],
        "snk_frames_create": [P(vp), i32, i64, i32, P(flt), P(flt), P(flt), P(dbl), P(dbl), i64, P(i64), P(i64), P(i64)],
        "snk_frames_destroy": [vp],
        "snk_concat_magphase_epoch": [vp, P(i64), i64, i32, i32, P(dbl), i32, P(dbl), P(dbl), P(dbl), P(dbl), P(dbl), P(dbl)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    _lib = lib
    return lib


STD_FLOAT32 = 1

EXPORTED_SYMBOLS = ["snk_last_error", "snk_version", "snk_device_count", "snk_db_create", "snk_db_destroy",
                    "snk_db_info", "snk_db_set_weights", "snk_db_set_engine", "snk_db_counters", "snk_db_profile_enable",
                    "snk_db_profile_read", "snk_knn",
                    "snk_knn_dev", "snk_knn_finish", "snk_debug_tc_keys", "snk_debug_greedy_one_keys", "snk_debug_greedy_one_times", "snk_debug_greedy_one_cta_times", "snk_topk_merge_dev", "snk_comm_unique_id", "snk_comm_init",
                    "snk_comm_info", "snk_comm_peer_exchange", "snk_knn_sharded_dev", "snk_knn_sharded_finish", "snk_greedy_batch",
                    "snk_greedy_batch_dev", "snk_greedy_batch_finish", "snk_greedy_sharded_batch_dev",
                    "snk_db_set_standardisation", "snk_prepare_targets", "snk_halfphone_targets",
                    "snk_halfphone_targets_dev", "snk_greedy_batch_unnorm",
                    "snk_greedy_batch_unnorm_dev",
                    "snk_candidate_distances", "snk_join_tiles", "snk_join_stats", "snk_join_viterbi_batch",
                    "snk_join_viterbi_batch_dev", "snk_acoustic_viterbi_batch", "snk_acoustic_viterbi_batch_dev",
                    "snk_acoustic_viterbi_finish", "snk_greedy_path_scores", "snk_frames_create", "snk_frames_destroy",
                    "snk_concat_magphase_epoch"]


def _check(rc):
    if rc != 0:
        raise EngineError(load_library().snk_last_error().decode("utf-8", "replace"))


def device_count():
    return load_library().snk_device_count()


UNIQUE_ID_BYTES = 128


def comm_unique_id():
    """The NCCL unique id (bytes) rank 0 hands to the other ranks before UnitDatabase.comm_init."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(load_library().snk_comm_unique_id(buf, UNIQUE_ID_BYTES))
    return buf.raw


class UnitDatabase:
    """A unit database resident in one GPU's HBM (F [N,Dt] f32, Jc [N+1,Dj] f32, unweighted)."""

    def __init__(self, F, Jc, multiepoch=1, layout=LAYOUT_SIMPLE, device=0):
        lib = load_library()
        F = np.ascontiguousarray(F, dtype=np.float32)
        Jc = np.ascontiguousarray(Jc, dtype=np.float32)
        if F.ndim != 2 or Jc.ndim != 2 or Jc.shape[0] != F.shape[0] + 1:
            raise ValueError("expected F [N,Dt] and Jc [N+1,Dj], got %s and %s" % (F.shape, Jc.shape))
        self.N, self.Dt = F.shape
        self.Dj = Jc.shape[1]
        self.multiepoch = int(multiepoch)
        self.device = int(device)
        self._h = C.c_void_p()
        _check(lib.snk_db_create(C.byref(self._h), self.device, self.N, self.Dt, self.Dj, self.multiepoch,
                                 _ptr(F, C.c_float), _ptr(Jc, C.c_float), layout))
        n, npr, dt, dj, m, jd = C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib.snk_db_info(self._h, C.byref(n), C.byref(npr), C.byref(dt), C.byref(dj), C.byref(m), C.byref(jd)))
        self.Nprime, self.joint_dim = npr.value, jd.value
        self.Djq = self.joint_dim - self.multiepoch * self.Dt

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load_library().snk_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- configuration
    def set_weights(self, wt, wj):
        wt = np.ascontiguousarray(wt, dtype=np.float64).ravel()
        wj = np.ascontiguousarray(wj, dtype=np.float64).ravel()
        if wt.size != self.Dt or wj.size != self.Dj:
            raise ValueError("weight vectors must have %d and %d entries" % (self.Dt, self.Dj))
        _check(load_library().snk_db_set_weights(self._h, _ptr(wt, C.c_double), _ptr(wj, C.c_double)))

    def set_engine(self, engine):
        _check(load_library().snk_db_set_engine(self._h, int(engine)))

    def counters(self, reset=False):
        out = np.zeros(4, dtype=np.int64)
        _check(load_library().snk_db_counters(self._h, _ptr(out, C.c_int64), int(reset)))
        return {"queries": int(out[0]), "recertified": int(out[1]), "launches": int(out[2]), "exhaustive": int(out[3])}

    def profile_enable(self, on=True):
        _check(load_library().snk_db_profile_enable(self._h, int(on)))

    def profile_read(self, which, reset=True):
        ms, n, w = C.c_double(), C.c_int64(), C.c_double()
        _check(load_library().snk_db_profile_read(self._h, int(which), C.byref(ms), C.byref(n), C.byref(w), int(reset)))
        return {"ms": ms.value, "launches": n.value, "work": w.value}

    # -- device-pointer entry points (used by distributed.py / bench.py with torch tensors: only raw addresses cross)
    def knn_dev(self, q_ptr, nq, k, dist_ptr, idx_ptr, space=SPACE_TARGET, id_offset=0, stream=0):
        _check(load_library().snk_knn_dev(self._h, space, C.c_void_p(q_ptr), int(nq), int(k), C.c_void_p(dist_ptr),
                                          C.c_void_p(idx_ptr), int(id_offset), C.c_void_p(stream)))

    def knn_finish(self):
        _check(load_library().snk_knn_finish(self._h))

    def greedy_batch_dev(self, targets_ptr, lens, paths_ptr, dists_ptr=0, start_states=None, stream=0, unnorm=False):
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        ss = None if start_states is None else np.ascontiguousarray(start_states, dtype=np.int64)
        lib = load_library()
        fn = lib.snk_greedy_batch_unnorm_dev if unnorm else lib.snk_greedy_batch_dev
        _check(fn(self._h, C.c_void_p(targets_ptr), _ptr(lens, C.c_int64), lens.size, _ptr(ss, C.c_int64),
                  C.c_void_p(paths_ptr), C.c_void_p(dists_ptr) if dists_ptr else None, C.c_void_p(stream)))

    def greedy_batch_finish(self):
        _check(load_library().snk_greedy_batch_finish(self._h))

    def greedy_sharded_batch_dev(self, targets_ptr, lens, jc_full_ptr, rows_full, id_offset, paths_ptr, dists_ptr=0,
                                 start_states=None, stream=0):
        """Collective: this handle holds joint rows [id_offset, id_offset + rows); jc_full_ptr is the replicated
        float32 join matrix on this device.  Follow with greedy_batch_finish() on every rank."""
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        ss = None if start_states is None else np.ascontiguousarray(start_states, dtype=np.int64)
        _check(load_library().snk_greedy_sharded_batch_dev(
            self._h, C.c_void_p(targets_ptr), _ptr(lens, C.c_int64), lens.size, _ptr(ss, C.c_int64), C.c_void_p(jc_full_ptr),
            int(rows_full), int(id_offset), C.c_void_p(paths_ptr), C.c_void_p(dists_ptr) if dists_ptr else None,
            C.c_void_p(stream)))

    def join_viterbi_batch_dev(self, cand_ptr, tdist_ptr, lens, K, paths_ptr, plen_ptr, pcost_ptr, tcost_ptr=0, jcost_ptr=0,
                               flags=0, stream=0):
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        _check(load_library().snk_join_viterbi_batch_dev(
            self._h, C.c_void_p(cand_ptr), C.c_void_p(tdist_ptr), _ptr(lens, C.c_int64), lens.size, int(K), int(flags),
            C.c_void_p(paths_ptr), C.c_void_p(plen_ptr), C.c_void_p(pcost_ptr), C.c_void_p(tcost_ptr) if tcost_ptr else None,
            C.c_void_p(jcost_ptr) if jcost_ptr else None, C.c_void_p(stream)))

    # -- database sharded over the GPUs of one box: NCCL communicator inside the library
    def comm_init(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), UNIQUE_ID_BYTES)
        _check(load_library().snk_comm_init(self._h, buf, int(rank), int(nranks)))
        self.comm_rank, self.comm_nranks = int(rank), int(nranks)

    def comm_info(self):
        r, n, v = C.c_int(), C.c_int(), C.c_int()
        _check(load_library().snk_comm_info(self._h, C.byref(r), C.byref(n), C.byref(v)))
        return {"rank": r.value, "nranks": n.value, "nccl_version": v.value,
                "peer_exchange": bool(load_library().snk_comm_peer_exchange(self._h))}

    def knn_sharded_dev(self, q_ptr, nq, k, dist_ptr, idx_ptr, space=SPACE_TARGET, id_offset=0, stream=0):
        _check(load_library().snk_knn_sharded_dev(self._h, space, C.c_void_p(q_ptr), int(nq), int(k), C.c_void_p(dist_ptr),
                                                  C.c_void_p(idx_ptr), int(id_offset), C.c_void_p(stream)))

    def knn_sharded_finish(self):
        _check(load_library().snk_knn_sharded_finish(self._h))

    def debug_tc_keys(self, Q, row0, nrows, space=SPACE_TARGET):
        """Raw tensor-core keys [nq, nrows] + the certificate's query norms, slack and max row norm (test instrumentation)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        keys = np.empty((Q.shape[0], nrows), dtype=np.float32)
        qn = np.empty(Q.shape[0], dtype=np.float32)
        eps, mx = C.c_float(), C.c_float()
        _check(load_library().snk_debug_tc_keys(self._h, space, _ptr(Q, C.c_double), Q.shape[0], int(row0), int(nrows),
                                                _ptr(keys, C.c_float), _ptr(qn, C.c_float), C.byref(eps), C.byref(mx)))
        return keys, qn, eps.value, mx.value

    def debug_greedy_one_keys(self, window, start_state=-1):
        """Keys [N'] of the single-utterance greedy kernel for the first step of an utterance starting with `window`
        [multiepoch, Dt] (weighted float64) + the query norm, slack and max row norm (test instrumentation)."""
        window = np.ascontiguousarray(window, dtype=np.float64)
        if window.shape != (self.multiepoch, self.Dt):
            raise ValueError("window must be [%d, %d], got %s" % (self.multiepoch, self.Dt, window.shape))
        keys = np.empty(self.Nprime, dtype=np.float32)
        qn, eps, mx = C.c_float(), C.c_float(), C.c_float()
        _check(load_library().snk_debug_greedy_one_keys(self._h, _ptr(window, C.c_double), int(start_state), _ptr(keys, C.c_float),
                                                        C.byref(qn), C.byref(eps), C.byref(mx)))
        return keys, qn.value, eps.value, mx.value

    # -- searches (host arrays in, host arrays out)
    def knn(self, Q, k, space=SPACE_TARGET):
        D = self.Dt if space == SPACE_TARGET else self.joint_dim
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.ndim != 2 or Q.shape[1] != D:
            raise ValueError("queries must be [nq, %d], got %s" % (D, Q.shape))
        nq = Q.shape[0]
        dist = np.empty((nq, k), dtype=np.float64)
        idx = np.empty((nq, k), dtype=np.int64)
        _check(load_library().snk_knn(self._h, space, _ptr(Q, C.c_double), nq, int(k), _ptr(dist, C.c_double),
                                      _ptr(idx, C.c_int64)))
        return dist, idx

    def set_standardisation(self, mean, std, special_uv_value=-1000.0, uv_scaling_factor=20.0):
        """mean_vec_target / std_vec_target of the voice + the constants of const.py:12-14 (row N4).
        float32 statistics (what the voice file holds) select numpy's float32 arithmetic, anything else float64."""
        f32 = np.asarray(mean).dtype == np.float32 and np.asarray(std).dtype == np.float32
        mean = np.ascontiguousarray(np.asarray(mean, dtype=np.float64).reshape(-1))
        std = np.ascontiguousarray(np.asarray(std, dtype=np.float64).reshape(-1))
        if mean.size != self.Dt or std.size != self.Dt:
            raise ValueError("mean / std must have %d entries" % self.Dt)
        _check(load_library().snk_db_set_standardisation(self._h, _ptr(mean, C.c_double), _ptr(std, C.c_double),
                                                         float(special_uv_value), float(uv_scaling_factor),
                                                         STD_FLOAT32 if f32 else 0))

    def prepare_targets(self, unnorm):
        """weight(standardise(unnorm)) on the device: float32 [T, Dt] -> float64 [T, Dt]."""
        unnorm = np.ascontiguousarray(unnorm, dtype=np.float32)
        if unnorm.ndim != 2 or unnorm.shape[1] != self.Dt:
            raise ValueError("unnorm speech must be [T, %d]" % self.Dt)
        out = np.empty(unnorm.shape, dtype=np.float64)
        _check(load_library().snk_prepare_targets(self._h, _ptr(unnorm, C.c_float), unnorm.shape[0],
                                                  _ptr(out, C.c_double)))
        return out

    def halfphone_targets(self, unnorm, points, durations=None):
        """weight(hstack(standardise(unnorm)[points], durations)) on the device: float32 [frames, dim], int [n, P]
        (+ float64 [n]) -> float64 [n, Dt]."""
        unnorm = np.ascontiguousarray(unnorm, dtype=np.float32)
        points = np.ascontiguousarray(points, dtype=np.int64)
        if unnorm.ndim != 2 or points.ndim != 2:
            raise ValueError("unnorm must be [frames, dim] and points [n, P]")
        dur = None if durations is None else np.ascontiguousarray(np.asarray(durations, dtype=np.float64).reshape(-1))
        if dur is not None and dur.shape[0] != points.shape[0]:
            raise ValueError("one duration per unit")
        out = np.empty((points.shape[0], self.Dt), dtype=np.float64)
        _check(load_library().snk_halfphone_targets(self._h, _ptr(unnorm, C.c_float), unnorm.shape[0], unnorm.shape[1],
                                                    _ptr(points, C.c_int64), points.shape[0], points.shape[1],
                                                    _ptr(dur, C.c_double), _ptr(out, C.c_double)))
        return out

    def greedy_batch_cat(self, cat, lens, start_states=None, return_dists=False, unnorm=False, as_arrays=False):
        """Batch given as one concatenated array [sum T_b, Dt] (may be pinned) + lengths: weighted float64
        unit features, or (unnorm=True) un-normalised float32 speech that the device standardises and weights.
        Returns one list of unit ids per utterance, as the reference does -- or, with as_arrays, one int64 array per
        utterance (views of one buffer): boxing 10^5 ids into Python ints costs several milliseconds per batch."""
        m = self.multiepoch
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        B = lens.size
        if B == 0:
            return ([], []) if return_dists else []
        if np.any(lens < m):
            raise ValueError("Not enough data points to segment array in 'cut' mode")  # segmentaxis.py:94-96
        cat = np.ascontiguousarray(cat, dtype=np.float32 if unnorm else np.float64)
        if cat.ndim != 2 or cat.shape[1] != self.Dt or cat.shape[0] != int(lens.sum()):
            raise ValueError("targets must be [sum(lens), %d]" % self.Dt)
        steps = lens // m
        paths = np.empty(int(steps.sum()), dtype=np.int64)
        dists = np.empty(int(steps.sum()), dtype=np.float64) if return_dists else None
        ss = None
        if start_states is not None:
            ss = np.ascontiguousarray(start_states, dtype=np.int64)
        lib = load_library()
        fn, ct = (lib.snk_greedy_batch_unnorm, C.c_float) if unnorm else (lib.snk_greedy_batch, C.c_double)
        _check(fn(self._h, _ptr(cat, ct), _ptr(lens, C.c_int64), B, _ptr(ss, C.c_int64), _ptr(paths, C.c_int64),
                  _ptr(dists, C.c_double)))
        cuts = np.cumsum(steps)[:-1]
        p = np.split(paths, cuts) if as_arrays else [x.tolist() for x in np.split(paths, cuts)]
        if return_dists:
            return p, np.split(dists, cuts)
        return p

    def greedy_batch(self, targets_list, start_states=None, return_dists=False, unnorm=False):
        for t in targets_list:
            if t.ndim != 2 or t.shape[1] != self.Dt:
                raise ValueError("each target utterance must be [T, %d]" % self.Dt)
        if len(targets_list) == 0:
            return ([], []) if return_dists else []
        lens = np.array([t.shape[0] for t in targets_list], dtype=np.int64)
        cat = np.concatenate(targets_list, axis=0)
        return self.greedy_batch_cat(cat, lens, start_states, return_dists, unnorm=unnorm)

    def candidate_distances(self, cand, targets):
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        T, K = cand.shape
        if targets.shape != (T, self.Dt):
            raise ValueError("targets must be [%d, %d]" % (T, self.Dt))
        dist = np.empty((T, K), dtype=np.float64)
        _check(load_library().snk_candidate_distances(self._h, _ptr(cand, C.c_int64), _ptr(targets, C.c_double), T, K,
                                                      _ptr(dist, C.c_double)))
        return dist

    def join_tiles(self, cand_list):
        K = cand_list[0].shape[1]
        lens = np.array([c.shape[0] for c in cand_list], dtype=np.int64)
        cat = np.ascontiguousarray(np.concatenate(cand_list, axis=0), dtype=np.int64)
        ntiles = int(np.maximum(lens - 1, 0).sum())
        tiles = np.empty((ntiles, K, K), dtype=np.float32)
        _check(load_library().snk_join_tiles(self._h, _ptr(cat, C.c_int64), _ptr(lens, C.c_int64), len(cand_list), K,
                                             _ptr(tiles, C.c_float)))
        return tiles

    def join_stats(self):
        """(finite entries, entries recomputed by direct differences) of the last join_tiles call (tensor-core path)."""
        out = np.zeros(2, dtype=np.int64)
        _check(load_library().snk_join_stats(self._h, _ptr(out, C.c_int64)))
        return int(out[0]), int(out[1])

    def join_viterbi_batch(self, cand_list, dist_list, flags=0, as_arrays=False):
        B = len(cand_list)
        if B == 0:
            return [], np.zeros(0), np.zeros(0), np.zeros(0)
        K = cand_list[0].shape[1]
        lens = np.array([c.shape[0] for c in cand_list], dtype=np.int64)
        cand = np.ascontiguousarray(np.concatenate(cand_list, axis=0), dtype=np.int64).reshape(-1, K)
        dist = np.ascontiguousarray(np.concatenate(dist_list, axis=0), dtype=np.float64).reshape(-1, K)
        if cand.shape != dist.shape:
            raise ValueError("candidates and distances differ in shape")
        paths = np.empty(max(int(lens.sum()), 1), dtype=np.int64)
        plen = np.empty(B, dtype=np.int64)
        pcost, tcost, jcost = (np.empty(B, dtype=np.float64) for _ in range(3))
        _check(load_library().snk_join_viterbi_batch(self._h, _ptr(cand, C.c_int64), _ptr(dist, C.c_double),
                                                     _ptr(lens, C.c_int64), B, K, flags, _ptr(paths, C.c_int64),
                                                     _ptr(plen, C.c_int64), _ptr(pcost, C.c_double),
                                                     _ptr(tcost, C.c_double), _ptr(jcost, C.c_double)))
        out, off = [], 0
        for b in range(B):
            seg = paths[off:off + max(int(plen[b]), 0)]
            out.append(seg if as_arrays else seg.tolist())
            off += lens[b]
        return out, pcost, tcost, jcost

    def acoustic_viterbi_batch_cat(self, cat, lens, K, flags=0, as_arrays=False):
        """preselect_units_acoustic + viterbi_search for a batch given as one concatenated float64 array [sum T_b, Dt]
        (may be pinned) + lengths; the candidate lists stay on the device.  Paths come back as lists of unit ids (the
        reference's form) or, with as_arrays, as int64 array views (no per-id boxing)."""
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        B = lens.size
        cat = np.ascontiguousarray(cat, dtype=np.float64)
        if cat.ndim != 2 or cat.shape[1] != self.Dt or cat.shape[0] != int(lens.sum()):
            raise ValueError("targets must be [sum(lens), %d]" % self.Dt)
        paths = np.empty(max(int(lens.sum()), 1), dtype=np.int64)
        plen = np.empty(B, dtype=np.int64)
        pcost, tcost, jcost = (np.empty(B, dtype=np.float64) for _ in range(3))
        _check(load_library().snk_acoustic_viterbi_batch(self._h, _ptr(cat, C.c_double), _ptr(lens, C.c_int64), B, int(K),
                                                         flags, _ptr(paths, C.c_int64), _ptr(plen, C.c_int64),
                                                         _ptr(pcost, C.c_double), _ptr(tcost, C.c_double),
                                                         _ptr(jcost, C.c_double)))
        out, off = [], 0
        for b in range(B):
            seg = paths[off:off + max(int(plen[b]), 0)]
            out.append(seg if as_arrays else seg.tolist())
            off += lens[b]
        return out, pcost, tcost, jcost

    def greedy_path_scores(self, targets, path, twidths, jwidths):
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        path = np.ascontiguousarray(path, dtype=np.int64)
        tw = np.ascontiguousarray(twidths, dtype=np.int32)
        jw = np.ascontiguousarray(jwidths, dtype=np.int32)
        P = path.size
        ts = np.empty((P, tw.size), dtype=np.float64)
        js = np.empty((max(P - 1, 0), jw.size), dtype=np.float64)
        _check(load_library().snk_greedy_path_scores(self._h, _ptr(targets, C.c_double), targets.shape[0],
                                                     _ptr(path, C.c_int64), P, _ptr(tw, C.c_int32), tw.size,
                                                     _ptr(jw, C.c_int32), jw.size, _ptr(ts, C.c_double),
                                                     _ptr(js, C.c_double)))
        return ts, js


class FrameStore:
    """Full-band MagPhase frames of the voice resident in HBM (row N2): gather + cross-fade + overlap-add of
    the selected units, as concatenateMagPhaseEpoch_sep_files does before waveform synthesis
    (reference script/synth_simple.py:677-747)."""

    def __init__(self, mag, real, imag, f0_interp, vuv, unit_frame, sent_lo, sent_hi, device=0):
        lib = load_library()
        mag, real, imag = (np.ascontiguousarray(a, dtype=np.float32) for a in (mag, real, imag))
        f0 = np.ascontiguousarray(f0_interp, dtype=np.float64).ravel()
        vv = np.ascontiguousarray(vuv, dtype=np.float64).ravel()
        uf, lo, hi = (np.ascontiguousarray(a, dtype=np.int64).ravel() for a in (unit_frame, sent_lo, sent_hi))
        if not (mag.shape == real.shape == imag.shape) or mag.ndim != 2 or f0.size != mag.shape[0] or vv.size != mag.shape[0]:
            raise ValueError("mag/real/imag must be [nframes, width] with f0 / vuv of nframes entries")
        if not (uf.size == lo.size == hi.size):
            raise ValueError("unit_frame, sent_lo, sent_hi must have one entry per unit")
        self.nframes, self.width = mag.shape
        self._h = C.c_void_p()
        _check(lib.snk_frames_create(C.byref(self._h), int(device), self.nframes, self.width, _ptr(mag, C.c_float),
                                     _ptr(real, C.c_float), _ptr(imag, C.c_float), _ptr(f0, C.c_double), _ptr(vv, C.c_double),
                                     uf.size, _ptr(uf, C.c_int64), _ptr(lo, C.c_int64), _ptr(hi, C.c_int64)))
        self.last_kernel_ms = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load_library().snk_frames_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def concatenate(self, path, multiepoch=1, overlap=0, fzero=None):
        """Returns (mag, real, imag, fz) float64 arrays ready for magphase.synthesis_from_lossless."""
        path = np.ascontiguousarray(path, dtype=np.int64)
        P = path.size
        if overlap % 2:
            raise AssertionError("frame overlap should be even number")
        taper = np.hanning(((overlap + 1) * 2) + 1)[1:overlap + 1].astype(np.float64) if overlap else None  # matrix_operations.py:19
        rows = P * int(multiepoch)
        mag, real, imag = (np.empty((rows, self.width)) for _ in range(3))
        fz, vuv = np.empty((rows, 1)), np.empty((rows, 1))
        ms = C.c_double()
        _check(load_library().snk_concat_magphase_epoch(self._h, _ptr(path, C.c_int64), P, int(multiepoch), int(overlap),
                                                        _ptr(taper, C.c_double), int(fzero is not None), _ptr(mag, C.c_double),
                                                        _ptr(real, C.c_double), _ptr(imag, C.c_double), _ptr(fz, C.c_double),
                                                        _ptr(vuv, C.c_double), C.byref(ms)))
        self.last_kernel_ms = ms.value
        if fzero is not None and np.size(fzero) > 0:
            fz = np.asarray(fzero)               # synth_simple.py:725-726
        return mag, real, imag, fz
