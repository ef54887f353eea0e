"""Drop-in for the KD-tree objects on snickery's search path.

`GpuKDTree` quacks like `scipy.spatial.cKDTree` for the calls the reference makes
(script/synth_simple.py:229,490; script/synth_halfphone.py:379,399,605,1364,1384,1932):
constructor from a data matrix, `.query(x, k=1, eps=0.0)` with scipy's shape conventions.
`GpuStashableKDTree` follows `sklearn.neighbors.KDTree.query` as used through the
reference's StashableKDTree (script/StashableKDTree.py:7-102; script/synth_halfphone.py:1287-1295;
script/active_learning_join.py:198-202): always 2-D outputs.

There is no tree: the search is an exact brute-force scan of the whole matrix on the GPU
(tensor-core distance GEMM shortlist + float64 re-rank, or the fp32 SIMT kernel), so `eps`,
`leafsize`, `balanced_tree` and `compact_nodes` are accepted and ignored -- an exact answer
satisfies every (1+eps) guarantee.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import engine


class GpuKDTree:
    def __init__(self, data, leafsize=16, compact_nodes=True, copy_data=False, balanced_tree=True,
                 boxsize=None, device=0, _db=None, _space=engine.SPACE_TARGET):
        if boxsize is not None:
            raise NotImplementedError("periodic boxes are not used by snickery and are not supported")
        if _db is not None:   # view over an existing resident database (used by Synthesiser)
            self._db, self._space = _db, _space
            self.n = _db.N if _space == engine.SPACE_TARGET else _db.Nprime
            self.m = _db.Dt if _space == engine.SPACE_TARGET else _db.joint_dim
            self.data = None
            return
        data = np.asarray(data)
        if data.ndim != 2:
            raise ValueError("data must be of shape (n, m)")
        self.n, self.m = data.shape
        # the engine stores float32 rows times float64 weights (the reference's own representation,
        # speech_manip.py:209-213); a generic float64 matrix is held as float32 x 1.0
        f32 = np.ascontiguousarray(data, dtype=np.float32)
        self.exact_storage = bool(np.array_equal(f32.astype(np.float64), np.asarray(data, dtype=np.float64)))
        if not self.exact_storage:
            # e.g. a stash of already weighted float64 rows (float32 voice * float64 weights): the products are not
            # float32 numbers.  Distances are then taken to the float32-rounded rows (each coordinate off by at most
            # 6e-8 relative), so neighbours tied to within that may come back in another order than scipy / sklearn
            # return them.  GpuKDTree.from_weighted(raw_f32, weights) keeps the exact values.
            warnings.warn("GpuKDTree: the data matrix is not exactly representable in float32; searching its float32 "
                          "rounding (use GpuKDTree.from_weighted(raw_float32, weight_vector) for exact float64 rows)",
                          RuntimeWarning, stacklevel=2)
        self.data = data
        self._db = engine.UnitDatabase(f32, np.zeros((self.n + 1, 1), np.float32), multiepoch=1, device=device)
        self._db.set_weights(np.ones(self.m), np.ones(1))
        self._space = engine.SPACE_TARGET

    @classmethod
    def from_weighted(cls, raw_f32, weight_vector, device=0):
        """Tree over weight(raw_f32, weight_vector) with the reference's exact float64 row values."""
        raw = np.ascontiguousarray(raw_f32, dtype=np.float32)
        self = cls.__new__(cls)
        self.n, self.m = raw.shape
        self.data = None
        self.exact_storage = True
        self._db = engine.UnitDatabase(raw, np.zeros((self.n + 1, 1), np.float32), multiepoch=1, device=device)
        self._db.set_weights(np.asarray(weight_vector, dtype=np.float64), np.ones(1))
        self._space = engine.SPACE_TARGET
        return self

    def query(self, x, k=1, eps=0.0, p=2, distance_upper_bound=np.inf, workers=1, n_jobs=None):
        if p != 2:
            raise NotImplementedError("only the Euclidean metric is used by snickery")
        if not np.isinf(distance_upper_bound):
            raise NotImplementedError("distance_upper_bound is not used by snickery")
        if eps < 0:
            raise ValueError("eps must be non-negative")
        x = np.asarray(x, dtype=np.float64)
        if x.shape[-1] != self.m:
            raise ValueError("x must consist of vectors of length %d but has shape %s" % (self.m, x.shape))
        single = x.ndim == 1
        lead = x.shape[:-1]
        q = x.reshape(-1, self.m)
        if isinstance(k, (list, tuple, np.ndarray)):
            ks = np.asarray(k, dtype=int)
            d, i = self._db.knn(q, int(ks.max()), self._space)
            d, i = d[:, ks - 1], i[:, ks - 1]
            return d.reshape(lead + (len(ks),)), i.reshape(lead + (len(ks),))
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        d, i = self._db.knn(q, k, self._space)
        if k == 1:
            d, i = d[:, 0], i[:, 0]
            if single:
                return float(d[0]), int(i[0])
            return d.reshape(lead), i.reshape(lead)
        return d.reshape(lead + (k,)), i.reshape(lead + (k,))


class GpuStashableKDTree(GpuKDTree):
    """sklearn-style: KDTree(X, leaf_size=100, metric='euclidean').query(X, k) -> 2-D (dist, ind)."""

    def __init__(self, data, leaf_size=40, metric="euclidean", device=0, **kwargs):
        if metric not in ("euclidean", "minkowski", "l2"):
            raise NotImplementedError("StashableKDTree is Euclidean only (StashableKDTree.py:86)")
        self.leaf_size = int(leaf_size)
        super().__init__(data, device=device)

    def query(self, X, k=1, return_distance=True, dualtree=False, breadth_first=False, sort_results=True):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("query data dimension must match training data dimension")
        d, i = self._db.knn(X, int(k), self._space)
        return (d, i) if return_distance else i

    # ---- persistence (StashableKDTree.py:43-83)
    def save_hdf(self, fname, interoperable=None):
        """Stash the tree in the reference's HDF5 layout (StashableKDTree.py:43-54): ``state_0..3`` = the arrays of
        sklearn's ``KDTree.__getstate__()`` (data, index array, node array, node bounds), ``int_values`` = its seven
        integers.  This engine has no tree, so for a stash the reference's own ``resurrect_tree`` can load the node arrays
        are built with scikit-learn on the host (``interoperable=True``; the default when scikit-learn is importable);
        without it ``state_1..3`` are written empty and the file round-trips through this class only."""
        if self.data is None:
            raise ValueError("this tree is a view over a resident database and holds no data matrix of its own")
        write_stash(fname, np.ascontiguousarray(self.data, dtype=np.float64), self.leaf_size, interoperable)

    @classmethod
    def load_hdf(cls, fname, device=0):
        """Rebuild from a stash: only ``state_0`` (the data matrix, StashableKDTree.py:15) is needed -- files
        written by the reference's sklearn-backed class load as well (their node arrays are skipped)."""
        from .hdf5_voice import Hdf5File
        f = Hdf5File(fname)
        if "state_0" not in f:
            raise ValueError("%s is not a StashableKDTree stash (no state_0)" % fname)
        return cls(f["state_0"], device=device)


def write_stash(fname, data, leaf_size=40, interoperable=None):
    """The HDF5 stash of StashableKDTree.save_hdf (StashableKDTree.py:43-54) for a data matrix.  interoperable: build
    sklearn's tree arrays so that the reference's sklearn-backed class can ``load_hdf`` the file (None: if scikit-learn is
    importable)."""
    from .hdf5_voice import save_voice
    state = None
    if interoperable is None or interoperable:
        try:
            from sklearn.neighbors import KDTree
            state = KDTree(data, leaf_size=int(leaf_size), metric="euclidean").__getstate__()
        except ImportError:
            if interoperable:
                raise
    if state is not None:
        arrays = {"state_%d" % i: np.ascontiguousarray(state[i]) for i in range(4)}
        arrays["int_values"] = np.asarray(state[4:11], dtype=np.int64)     # leaf_size, n_levels, n_nodes, n_trims, n_leaves, n_splits, n_calls
    else:
        arrays = {"state_0": data, "state_1": np.zeros((0,), np.int64), "state_2": np.zeros((0,), np.float64),
                  "state_3": np.zeros((0,), np.float64),
                  "int_values": np.array([int(leaf_size), 0, 0, 0, 0, 0, 0], dtype=np.int64)}
    save_voice(fname, arrays, chunked=())


def resurrect_tree(fname, device=0):
    """StashableKDTree.py:97-102."""
    return GpuStashableKDTree.load_hdf(fname, device=device)
