"""Row N3: the voice-file reader.  Pinned against a file written by the real HDF5 library (a MATLAB 7.3
MAT-file from scipy's test data, copied to tests/golden/libhdf5_testdouble.mat: user block, version-0
superblock, symbol-table root group, contiguous float64 dataset); the voice schema itself is exercised
through save_voice, which lays files out the way h5py's defaults do for train_simple.py:93-142."""
import os

import numpy as np
import pytest

from snickery_b200.hdf5_voice import Hdf5File, Hdf5FormatError, load_voice, save_voice

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def make_voice(N=1200, Dt=61, Dj=148, seed=0):
    rng = np.random.default_rng(seed)
    return {
        "train_unit_features": rng.normal(size=(N, Dt)).astype(np.float32),
        "join_contexts": rng.normal(size=(N + 1, Dj)).astype(np.float32),
        "mean_target": rng.normal(size=Dt).astype(np.float32),
        "std_target": rng.uniform(0.5, 2, size=Dt).astype(np.float32),
        "mean_join": rng.normal(size=Dj // 4).astype(np.float32),
        "std_join": rng.uniform(0.5, 2, size=Dj // 4).astype(np.float32),
        "train_unit_names": np.array([b"a/b/c_L/d/e_%d" % i for i in range(N)], dtype="S50"),
        "filenames": np.array([b"utt_%04d" % (i // 100) for i in range(N)], dtype="S50"),
        "unit_index_within_sentence_dset": (np.arange(N) % 100).astype(np.int32),
    }


def test_reads_file_written_by_libhdf5():
    f = Hdf5File(os.path.join(GOLDEN, "libhdf5_testdouble.mat"))
    assert f.keys() == ["testdouble"]
    info = f.info("testdouble")
    assert info.layout == "contiguous" and info.dtype == np.float64
    a = f["testdouble"]
    # scipy's expectation for this test variable: pi/4 * arange(9); MATLAB stores it column-major
    np.testing.assert_allclose(a.ravel(), np.pi / 4 * np.arange(9), rtol=0, atol=1e-15)


@pytest.mark.parametrize("kw", [dict(), dict(chunk_bytes=4096), dict(chunk_bytes=700, gzip=4, shuffle=True),
                                dict(chunked=())])
def test_voice_round_trip(tmp_path, kw):
    v = make_voice()
    path = str(tmp_path / "voice.hdf5")
    save_voice(path, v, **kw)
    f = Hdf5File(path)
    assert sorted(f.keys()) == sorted(v)
    w = load_voice(path)
    for k in v:
        assert w[k].dtype == v[k].dtype and w[k].shape == v[k].shape and np.array_equal(w[k], v[k]), k
    if kw.get("chunk_bytes") == 700:       # one row per chunk: a three-level chunk index
        assert f.info("join_contexts").chunk == (1, 148, 4) and len(f.info("join_contexts").filters) == 2
    # reading into caller-provided (e.g. pinned) memory
    out = np.empty(v["join_contexts"].shape, np.float32)
    assert f.read("join_contexts", out=out) is out and np.array_equal(out, v["join_contexts"])
    with pytest.raises(ValueError):
        f.read("join_contexts", out=np.empty((3, 3), np.float32))


def test_optional_magphase_and_empty(tmp_path):
    v = make_voice(N=50)
    v["mp_mag"] = np.random.default_rng(1).normal(size=(200, 33)).astype(np.float32)
    v["mp_fz"] = np.zeros((0,), np.float32)
    path = str(tmp_path / "v.hdf5")
    save_voice(path, v)
    w = load_voice(path)
    assert np.array_equal(w["mp_mag"], v["mp_mag"]) and w["mp_fz"].shape == (0,)
    assert "mp_mag" not in load_voice(path, optional=False)


def test_errors(tmp_path):
    p = str(tmp_path / "x.hdf5")
    with open(p, "wb") as f:
        f.write(b"not an hdf5 file" * 100)
    with pytest.raises(Hdf5FormatError, match="signature"):
        Hdf5File(p)
    with open(p, "wb") as f:      # libver='latest' superblock
        f.write(b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + b"\0" * 64)
    with pytest.raises(Hdf5FormatError, match="superblock version 2"):
        Hdf5File(p)
    v = make_voice(N=20)
    del v["join_contexts"]
    save_voice(p, v)
    with pytest.raises(Hdf5FormatError, match="join_contexts"):
        load_voice(p)
    v = make_voice(N=20)
    v["join_contexts"] = v["join_contexts"][:-1]
    save_voice(p, v)
    with pytest.raises(Hdf5FormatError, match=r"N \+ 1"):
        load_voice(p)


def test_stash_with_unsupported_member(tmp_path, monkeypatch):
    """A file may hold datasets of a type this reader does not know (here: a variable-length type); it must still open
    and serve the others, failing only on access.  (Compound types -- sklearn's node array in a StashableKDTree stash --
    are read: test_stash_written_for_the_reference_is_a_valid_sklearn_state.)"""
    import struct
    from snickery_b200 import hdf5_voice
    real = hdf5_voice._dtype_message

    def fake(dt):   # float32 members get a variable-length (class 9) datatype message
        if np.dtype(dt) == np.float32:
            return struct.pack("<BBBBI", 0x19, 0, 0, 0, 4)
        return real(dt)
    monkeypatch.setattr(hdf5_voice, "_dtype_message", fake)
    data = np.random.default_rng(3).normal(size=(40, 7))
    path = str(tmp_path / "stash.hdf5")
    save_voice(path, {"state_0": data, "state_2": np.zeros(9, np.float32), "int_values": np.arange(7)}, chunked=())
    f = Hdf5File(path)
    assert np.array_equal(f["state_0"], data) and np.array_equal(f["int_values"], np.arange(7))
    with pytest.raises(Hdf5FormatError, match="datatype class 9"):
        f["state_2"]


def test_round_trip_fuzz(tmp_path):
    """Random shapes, dtypes, chunk sizes and filter pipelines (seeded): what is written is what is read."""
    rng = np.random.default_rng(2024)
    for trial in range(25):
        arrays, chunked = {}, []
        for i in range(int(rng.integers(1, 8))):
            kind = rng.integers(0, 5)
            rows = int(rng.integers(0, 400))
            shape = (rows,) if rng.random() < 0.4 else (rows, int(rng.integers(1, 70)))
            if kind == 0:
                a = rng.normal(size=shape).astype(np.float32)
            elif kind == 1:
                a = rng.normal(size=shape)
            elif kind == 2:
                a = rng.integers(-2 ** 31, 2 ** 31 - 1, size=shape).astype(np.int32)
            elif kind == 3:
                a = rng.integers(0, 2 ** 62, size=shape).astype(np.int64)
            else:
                a = np.array([("u%d" % v).encode() for v in rng.integers(0, 10 ** 6, size=int(np.prod(shape)))] or [b""],
                             dtype="S%d" % int(rng.integers(8, 51)))[: int(np.prod(shape))].reshape(shape)
            name = "d%d_%s" % (i, "x" * int(rng.integers(0, 12)))
            arrays[name] = a
            if rng.random() < 0.6:
                chunked.append(name)
        path = str(tmp_path / ("f%d.hdf5" % trial))
        kw = {}
        if rng.random() < 0.5:
            kw["gzip"] = int(rng.integers(1, 9))
        if rng.random() < 0.5:
            kw["shuffle"] = True
        save_voice(path, arrays, chunked=tuple(chunked), chunk_bytes=int(rng.integers(64, 20000)), **kw)
        f = Hdf5File(path)
        assert sorted(f.keys()) == sorted(arrays)
        for k, a in arrays.items():
            b = f[k]
            assert b.shape == a.shape and b.dtype == a.dtype and np.array_equal(b, a), (trial, k, a.shape, a.dtype, kw)


def test_stash_written_for_the_reference_is_a_valid_sklearn_state(tmp_path):
    """StashableKDTree.save_hdf / load_hdf (reference script/StashableKDTree.py:43-83): state_0..3 are the arrays of sklearn's
    KDTree.__getstate__(), int_values its seven integers.  The stash this package writes must carry exactly those (the node
    array is a compound HDF5 type), so that the reference's sklearn-backed resurrect_tree can __setstate__ it: rebuild a
    sklearn KDTree from the file's arrays and query it."""
    sklearn_neighbors = pytest.importorskip("sklearn.neighbors")
    from snickery_b200.kdtree import write_stash
    rng = np.random.default_rng(7)
    X = rng.standard_normal((700, 9))
    fn = str(tmp_path / "stash.hdf5")
    write_stash(fn, X, leaf_size=25, interoperable=True)
    f = Hdf5File(fn)
    ref_tree = sklearn_neighbors.KDTree(X, leaf_size=25, metric="euclidean")
    ref = ref_tree.__getstate__()
    loaded = [f["state_%d" % i] for i in range(4)]
    assert loaded[2].dtype.names == ("idx_start", "idx_end", "is_leaf", "radius")
    for i in range(4):
        assert loaded[i].dtype == ref[i].dtype and loaded[i].shape == ref[i].shape
        assert loaded[i].tobytes() == np.ascontiguousarray(ref[i]).tobytes()
    ints = [int(v) for v in f["int_values"]]
    assert ints == [int(v) for v in ref[4:11]] and ints[0] == 25
    # what load_hdf does (StashableKDTree.py:56-80), with the metric object(s) sklearn's own state carries
    tree = sklearn_neighbors.KDTree.__new__(sklearn_neighbors.KDTree)
    tree.__setstate__(tuple(loaded + ints + list(ref[11:])))
    q = rng.standard_normal((30, 9))
    d1, i1 = tree.query(q, k=6)
    d2, i2 = ref_tree.query(q, k=6)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # without scikit-learn's arrays the file still round-trips through this package (state_0 is all it needs)
    fn2 = str(tmp_path / "plain.hdf5")
    write_stash(fn2, X, interoperable=False)
    assert np.array_equal(Hdf5File(fn2)["state_0"], X) and Hdf5File(fn2)["state_1"].size == 0


@pytest.mark.parametrize("version", [1, 2, 3])
def test_compound_datatype_versions_are_read(tmp_path, monkeypatch, version):
    """libhdf5 writes compound datatype messages in three encodings (version 1: padded names + a dimensionality block,
    version 2: padded names, version 3: unpadded names and an offset as wide as the element size needs).  The writer here
    emits version 1; the reader must take all three (a stash written with libver='latest' carries version 3)."""
    import struct
    from snickery_b200 import hdf5_voice
    real = hdf5_voice._dtype_message
    dt = np.dtype([("idx_start", "<i8"), ("idx_end", "<i8"), ("is_leaf", "<i8"), ("radius", "<f8")])

    def encode(d):
        d = np.dtype(d)
        if not d.names or version == 1:
            return real(d)
        body = struct.pack("<BBBBI", 0x06 | (version << 4), len(d.names), 0, 0, d.itemsize)
        for name in d.names:
            sub, off = d.fields[name][0], d.fields[name][1]
            nm = name.encode("ascii") + b"\0"
            if version == 2:
                nm += b"\0" * (-len(nm) % 8)
                body += nm + struct.pack("<I", off)
            else:
                body += nm + struct.pack("<B", off)              # element size 32 < 256: one byte
            body += hdf5_voice._atomic_dtype_message(sub)
        return body
    monkeypatch.setattr(hdf5_voice, "_dtype_message", encode)
    rng = np.random.default_rng(version)
    a = np.zeros(37, dtype=dt)
    a["idx_start"], a["idx_end"] = rng.integers(0, 1000, 37), rng.integers(0, 1000, 37)
    a["is_leaf"], a["radius"] = rng.integers(0, 2, 37), rng.random(37)
    path = str(tmp_path / "c.hdf5")
    save_voice(path, {"state_2": a, "other": np.arange(5.0)}, chunked=())
    f = Hdf5File(path)
    b = f["state_2"]
    assert b.dtype == dt and b.tobytes() == a.tobytes()
    assert np.array_equal(f["other"], np.arange(5.0))
