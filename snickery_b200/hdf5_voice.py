"""Voice-file front end: read the reference's ``.hdf5`` database dump without h5py.

The reference writes its voice with ``h5py.File(fname, "w")`` and plain ``create_dataset`` calls
(``train_simple.py:93-142``; ``train_halfphone.py`` does the same) and reads it back dataset by dataset
(``synth_simple.py:76-106``).  h5py's default ``libver`` produces the original HDF5 file layout:
version-0/1 superblock, version-1 object headers, symbol-table groups (local heap + version-1 B-tree +
``SNOD`` nodes), contiguous storage for fixed-shape datasets and chunked storage (version-1 chunk B-tree,
no filters) for the ones created with ``maxshape``.  This module reads exactly that subset -- plus
compact storage, the deflate / shuffle / fletcher32 filters and a user block -- with numpy only, so a
voice trained by the reference loads straight into the resident device layout.

It is a reader for this path's input format, not a general HDF5 library: new-style groups (fractal
heaps), version-2 object headers, variable-length and compound types raise ``Hdf5FormatError``.

Schema (``doc/content.tex:27-35``; ``train_simple.py:94-142``): ``train_unit_features`` f32 [N, Dt],
``join_contexts`` f32 [N+1, Dj], ``mean_target`` / ``std_target`` / ``mean_join`` / ``std_join`` f32,
``train_unit_names`` / ``filenames`` |S50 [N], ``unit_index_within_sentence_dset`` i32 [N] and, with
``store_full_magphase``, ``mp_mag`` / ``mp_imag`` / ``mp_real`` / ``mp_fz``.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Hdf5FormatError(ValueError):
    pass


class _Dataset:
    __slots__ = ("name", "shape", "maxshape", "dtype", "layout", "address", "size", "chunk", "filters", "compact")

    def __repr__(self):
        return "<dataset %r shape=%s dtype=%s %s>" % (self.name, self.shape, self.dtype, self.layout)


class Hdf5File:
    """Read-only view of the datasets in the root group (and nested old-style groups, '/'-joined)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self._buf = memoryview(f.read())
        self.path = path
        self._find_superblock()
        self.datasets = {}
        self._walk_group(self._root_btree, self._root_heap, "")

    # ---- low level
    def _u(self, off, n):
        return int.from_bytes(self._buf[off:off + n], "little")

    def _addr(self, off):
        a = self._u(off, self._O)
        return None if a == (1 << (8 * self._O)) - 1 else a + self._base

    def _find_superblock(self):
        buf, off = self._buf, 0
        while off + 8 <= len(buf):
            if bytes(buf[off:off + 8]) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2      # user block: 512, 1024, 2048, ...
        else:
            raise Hdf5FormatError("%s: no HDF5 signature" % self.path)
        ver = buf[off + 8]
        if ver > 1:
            raise Hdf5FormatError("superblock version %d (libver='latest' files) is not supported; "
                                  "the reference writes version 0" % ver)
        self._O, self._L = buf[off + 13], buf[off + 14]
        p = off + 24 + (4 if ver == 1 else 0)
        self._base = 0
        self._base = self._u(p, self._O)      # every other address in the file is relative to this one
        p += 4 * self._O
        # root group symbol table entry
        self._root_header = self._addr(p + self._O)
        cache = self._u(p + 2 * self._O, 4)
        if cache == 1:
            s = p + 2 * self._O + 8
            self._root_btree, self._root_heap = self._addr(s), self._addr(s + self._O)
        else:
            msgs = self._messages(self._root_header)
            st = [m for m in msgs if m[0] == 0x11]
            if not st:
                raise Hdf5FormatError("root group is not a symbol-table group")
            self._root_btree, self._root_heap = self._addr(st[0][1]), self._addr(st[0][1] + self._O)

    def _messages(self, addr):
        """[(type, data offset, size)] of a version-1 object header, following continuation blocks."""
        buf = self._buf
        if bytes(buf[addr:addr + 4]) == b"OHDR":
            raise Hdf5FormatError("version-2 object headers are not supported")
        if buf[addr] != 1:
            raise Hdf5FormatError("object header version %d" % buf[addr])
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                data = p + 8
                if mtype == 0x10:
                    blocks.append((self._addr(data), self._u(data + self._O, self._L)))
                out.append((mtype, data, msize))
                p = data + msize
        return out

    def _heap_name(self, heap, off):
        if bytes(self._buf[heap:heap + 4]) != b"HEAP":
            raise Hdf5FormatError("bad local heap")
        seg = self._addr(heap + 8 + 2 * self._L)
        p = seg + off
        end = p
        while self._buf[end] != 0:
            end += 1
        return bytes(self._buf[p:end]).decode("ascii", "replace")

    def _walk_group(self, btree, heap, prefix):
        for name, header, cache, scratch in self._group_entries(btree, heap):
            full = prefix + name
            msgs = self._messages(header)
            st = [m for m in msgs if m[0] == 0x11]
            if st:
                self._walk_group(self._addr(st[0][1]), self._addr(st[0][1] + self._O), full + "/")
            elif any(m[0] == 0x08 for m in msgs):
                try:
                    self.datasets[full] = self._parse_dataset(full, msgs)
                except Hdf5FormatError as e:          # unsupported member: report on access, not on open
                    self.datasets[full] = e

    def _group_entries(self, node, heap):
        buf = self._buf
        if node is None:
            return
        sig = bytes(buf[node:node + 4])
        if sig == b"TREE":
            if buf[node + 4] != 0:
                raise Hdf5FormatError("group B-tree expected")
            n = self._u(node + 6, 2)
            p = node + 8 + 2 * self._O
            for i in range(n):
                child = self._addr(p + self._L + i * (self._L + self._O))
                for e in self._group_entries(child, heap):
                    yield e
        elif sig == b"SNOD":
            n = self._u(node + 6, 2)
            esz = 2 * self._O + 24
            for i in range(n):
                p = node + 8 + i * esz
                yield (self._heap_name(heap, self._u(p, self._O)), self._addr(p + self._O),
                       self._u(p + 2 * self._O, 4), p + 2 * self._O + 8)
        else:
            raise Hdf5FormatError("bad group node signature %r" % sig)

    # ---- dataset description
    def _parse_dataset(self, name, msgs):
        buf = self._buf
        d = _Dataset()
        d.name, d.filters, d.chunk, d.compact, d.address, d.size = name, [], None, None, None, 0
        for mtype, p, size in msgs:
            if mtype == 0x01:                                   # dataspace
                ver, rank, flags = buf[p], buf[p + 1], buf[p + 2]
                q = p + (8 if ver == 1 else 4)
                d.shape = tuple(self._u(q + i * self._L, self._L) for i in range(rank))
                q += rank * self._L
                d.maxshape = tuple(self._u(q + i * self._L, self._L) for i in range(rank)) if flags & 1 else d.shape
            elif mtype == 0x03:                                 # datatype
                d.dtype = self._parse_dtype(p)
            elif mtype == 0x08:                                 # layout
                ver = buf[p]
                if ver == 3:
                    cls = buf[p + 1]
                    if cls == 0:
                        n = self._u(p + 2, 2)
                        d.layout, d.compact = "compact", (p + 4, n)
                    elif cls == 1:
                        d.layout, d.address, d.size = "contiguous", self._addr(p + 2), self._u(p + 2 + self._O, self._L)
                    elif cls == 2:
                        nd = buf[p + 2]
                        d.layout, d.address = "chunked", self._addr(p + 3)
                        d.chunk = tuple(self._u(p + 3 + self._O + 4 * i, 4) for i in range(nd))
                    else:
                        raise Hdf5FormatError("%s: layout class %d" % (name, cls))
                elif ver in (1, 2):
                    nd, cls = buf[p + 1], buf[p + 2]
                    q = p + 8
                    if cls != 0:
                        d.address = self._addr(q)
                        q += self._O
                    dims = tuple(self._u(q + 4 * i, 4) for i in range(nd))
                    q += 4 * nd
                    if cls == 2:
                        # as in version 3, the dimensionality of a chunked layout counts the trailing element-size
                        # dimension: the nd values just read are the whole chunk shape, nothing follows them
                        d.layout, d.chunk = "chunked", dims
                    elif cls == 1:
                        d.layout = "contiguous"
                    else:
                        n = self._u(q, 4)
                        d.layout, d.compact = "compact", (q + 4, n)
                else:
                    raise Hdf5FormatError("%s: layout message version %d" % (name, ver))
            elif mtype == 0x0B:                                 # filter pipeline
                ver, nf = buf[p], buf[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self._u(q, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self._u(q + 2, 2)
                        q += 4
                    else:
                        nlen = 0
                        q += 2
                    ncd = self._u(q + 2, 2)
                    q += 4
                    q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = [self._u(q + 4 * i, 4) for i in range(ncd)]
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    d.filters.append((fid, cd))
        if not hasattr(d, "shape") or not hasattr(d, "dtype") or not hasattr(d, "layout"):
            raise Hdf5FormatError("%s: incomplete dataset header" % name)
        return d

    def _parse_dtype(self, p):
        buf = self._buf
        cls, ver = buf[p] & 0x0F, buf[p] >> 4
        bits0 = buf[p + 1]
        size = self._u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            kind = "i" if bits0 & 0x08 else "u"
            return np.dtype("%s%s%d" % (order, kind, size))
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5FormatError("float of %d bytes" % size)
            return np.dtype("%sf%d" % (order, size))
        if cls == 3:
            return np.dtype("S%d" % size)
        if cls == 6:
            n = buf[p + 1] | (buf[p + 2] << 8)
            q = p + 8
            names, formats, offsets = [], [], []
            for _ in range(n):
                end = q
                while buf[end] != 0:
                    end += 1
                names.append(bytes(buf[q:end]).decode("ascii"))
                if ver < 3:
                    q += (end - q + 8) // 8 * 8            # name + NUL padded to a multiple of eight
                    offsets.append(self._u(q, 4))
                    q += 4
                    if ver == 1:
                        if buf[q] != 0:
                            raise Hdf5FormatError("array members of a compound are not supported")
                        q += 28                            # dimensionality, reserved, permutation, reserved, 4 sizes
                else:
                    q = end + 1
                    w = 1 if size < (1 << 8) else 2 if size < (1 << 16) else 4 if size < (1 << 32) else 8
                    offsets.append(self._u(q, w))
                    q += w
                sub = self._parse_dtype(q)
                if sub.names:
                    raise Hdf5FormatError("nested compounds are not supported")
                formats.append(sub)
                q += 8 + {"i": 4, "u": 4, "f": 12, "S": 0}[sub.kind]
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
        raise Hdf5FormatError("datatype class %d (version %d) is not supported" % (cls, ver))

    # ---- data
    def keys(self):
        return list(self.datasets)

    def __contains__(self, name):
        return name in self.datasets

    def info(self, name):
        d = self.datasets[name]
        if isinstance(d, Exception):
            raise d
        return d

    def __getitem__(self, name):
        return self.read(name)

    def read(self, name, out=None):
        """The whole dataset as a numpy array (native byte order).  out: optional preallocated array (for
        example pinned host memory) of the dataset's shape and dtype."""
        d = self.info(name)
        n = int(np.prod(d.shape, dtype=np.int64)) if d.shape else 1
        native = d.dtype.newbyteorder("=")
        if out is None:
            out = np.empty(d.shape, dtype=native)
        elif out.shape != d.shape or out.dtype != native:
            raise ValueError("out must be %s %s" % (d.shape, native))
        if n == 0:
            return out
        if d.layout == "compact":
            off, nbytes = d.compact
            out[...] = np.frombuffer(self._buf[off:off + nbytes], dtype=d.dtype, count=n).reshape(d.shape)
        elif d.layout == "contiguous":
            if d.address is None:
                out[...] = 0                # never written: the fill value (zero unless a fill message says otherwise)
            else:
                out[...] = np.frombuffer(self._buf[d.address:d.address + n * d.dtype.itemsize], dtype=d.dtype,
                                         count=n).reshape(d.shape)
        else:
            out[...] = 0
            if d.address is not None:
                self._read_chunks(d, d.address, out)
        return out

    def _read_chunks(self, d, node, out):
        buf = self._buf
        if bytes(buf[node:node + 4]) != b"TREE" or buf[node + 4] != 1:
            raise Hdf5FormatError("%s: bad chunk B-tree node" % d.name)
        level, n = buf[node + 5], self._u(node + 6, 2)
        nd = len(d.chunk)
        ksz = 8 + 8 * nd
        p = node + 8 + 2 * self._O
        cdims = d.chunk[:-1]
        for i in range(n):
            k = p + i * (ksz + self._O)
            child = self._addr(k + ksz)
            if level > 0:
                self._read_chunks(d, child, out)
                continue
            nbytes, mask = self._u(k, 4), self._u(k + 4, 4)
            offs = tuple(self._u(k + 8 + 8 * j, 8) for j in range(nd - 1))
            raw = bytes(buf[child:child + nbytes])
            for fi in range(len(d.filters) - 1, -1, -1):
                if mask >> fi & 1:
                    continue
                fid, cd = d.filters[fi]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else d.dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = a.size // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise Hdf5FormatError("%s: filter %d is not supported" % (d.name, fid))
            chunk = np.frombuffer(raw, dtype=d.dtype, count=int(np.prod(cdims))).reshape(cdims)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, d.shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]


VOICE_ARRAYS = ("train_unit_features", "join_contexts", "mean_target", "std_target", "mean_join", "std_join")
VOICE_OPTIONAL = ("train_unit_names", "filenames", "unit_index_within_sentence_dset", "cutpoints",
                  "mp_mag", "mp_imag", "mp_real", "mp_fz")


def load_voice(path, optional=True):
    """The arrays Synthesiser.__init__ reads from the database dump (synth_simple.py:89-106), same names,
    same dtypes (float32 matrices, byte-string labels)."""
    f = Hdf5File(path)
    missing = [k for k in VOICE_ARRAYS if k not in f]
    if missing:
        raise Hdf5FormatError("%s lacks %s (found %s)" % (path, missing, f.keys()))
    voice = {k: f[k] for k in VOICE_ARRAYS}
    if optional:
        for k in VOICE_OPTIONAL:
            if k in f:
                voice[k] = f[k]
    N = voice["train_unit_features"].shape[0]
    if voice["join_contexts"].shape[0] != N + 1:
        raise Hdf5FormatError("join_contexts has %d rows for %d units (expected N + 1, train_simple.py:142)"
                              % (voice["join_contexts"].shape[0], N))
    return voice


# ---------------------------------------------------------------------------------------------------
# Writer for the same subset.  The loader above is what the path needs; this exists so that synthetic
# voices (snickery_b200.synthetic) and the test fixtures can be put in the reference's container format
# on a machine without h5py.  It lays the file out the way h5py's defaults do for train_simple.py's
# calls: version-0 superblock, one symbol-table root group, contiguous storage, and chunked storage
# (version-1 chunk B-tree, two levels when needed) for the datasets named in `chunked`.
_UNDEF = (1 << 64) - 1
_LEAF_K, _INTERNAL_K, _CHUNK_K = 16, 16, 32


def _atomic_dtype_message(dt):
    """Datatype message of an atomic type: 8-byte header + properties, nothing after them (what a compound member embeds)."""
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, mant_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        head = struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, mant_bits, exp_bits, 0, mant_bits, bias)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    raise Hdf5FormatError("cannot write dtype %s" % dt)


def _dtype_message(dt):
    """Datatype message (type 0x03).  Structured dtypes become a version-1 compound (the form libhdf5 writes by default,
    e.g. for sklearn's KDTree node array in a StashableKDTree stash): per member the name padded to 8 bytes, byte offset,
    a scalar dimensionality block and the member's own atomic datatype message."""
    dt = np.dtype(dt)
    if dt.names:
        body = struct.pack("<BBBBI", 0x16, len(dt.names) & 0xff, len(dt.names) >> 8, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            if sub.names or sub.shape:
                raise Hdf5FormatError("nested / array members are not supported (%s)" % name)
            nm = name.encode("ascii") + b"\0"
            nm += b"\0" * (-len(nm) % 8)
            body += nm + struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0) + _atomic_dtype_message(sub)
        return body
    return _atomic_dtype_message(dt)


def _message(mtype, body):
    body += b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def save_voice(path, arrays, chunked=("train_unit_features", "join_contexts", "train_unit_names", "filenames",
                                      "unit_index_within_sentence_dset"), chunk_bytes=1 << 20, gzip=None,
               shuffle=False):
    """arrays: {name: ndarray}.  Datasets in `chunked` (the ones train_simple.py:135-142 creates with
    maxshape) are stored in row chunks of about chunk_bytes, everything else contiguously.  gzip (level) /
    shuffle add h5py's compression="gzip" / shuffle=True pipeline to the chunked datasets."""
    names = sorted(arrays)
    if len(names) > 2 * _LEAF_K:
        raise Hdf5FormatError("at most %d datasets" % (2 * _LEAF_K))
    out = bytearray(96)                      # superblock, filled in last

    def alloc(data):
        out.extend(b"\0" * (-len(out) % 8))
        at = len(out)
        out.extend(data)
        return at

    headers = {}
    for name in names:
        a = np.ascontiguousarray(arrays[name])
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        rank = a.ndim
        is_chunked = name in chunked and a.size > 0 and rank >= 1
        space = struct.pack("<BBB5x", 1, rank, 1 if is_chunked else 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        if is_chunked:
            space += b"".join(struct.pack("<Q", s) for s in a.shape)
        if not is_chunked:
            addr = alloc(a.tobytes()) if a.size else _UNDEF
            layout = struct.pack("<BBQQ", 3, 1, addr, a.nbytes)
        else:
            row_bytes = max(1, a.nbytes // a.shape[0])
            rows = int(max(1, min(a.shape[0], chunk_bytes // row_bytes)))
            cdims = (rows,) + a.shape[1:]
            entries = []
            for r0 in range(0, a.shape[0], rows):
                block = np.zeros(cdims, dtype=a.dtype)          # edge chunks are stored whole
                part = a[r0:r0 + rows]
                block[:part.shape[0]] = part
                raw = block.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, a.dtype.itemsize).T.tobytes()
                if gzip is not None:
                    raw = zlib.compress(raw, gzip)
                entries.append((r0, alloc(raw), len(raw)))

            def key(r0, nbytes):
                return struct.pack("<II", nbytes, 0) + struct.pack("<Q", r0) + b"\0" * (8 * rank)

            def node(level, items):          # items: (first row, child address, chunk bytes)
                body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), _UNDEF, _UNDEF)
                for r0, child, nbytes in items:
                    body += key(r0, nbytes) + struct.pack("<Q", child)
                body += key(a.shape[0] + (-a.shape[0] % rows), 0)
                full = 24 + 2 * _CHUNK_K * (16 + 8 * rank + 8) + (16 + 8 * rank)
                return alloc(body + b"\0" * max(0, full - len(body)))

            level = 0
            while len(entries) > 2 * _CHUNK_K:
                groups = [entries[i:i + 2 * _CHUNK_K] for i in range(0, len(entries), 2 * _CHUNK_K)]
                entries = [(g[0][0], node(level, g), g[0][2]) for g in groups]
                level += 1
            root = node(level, entries)
            layout = struct.pack("<BBBQ", 3, 2, rank + 1, root) + b"".join(struct.pack("<I", c) for c in cdims)
            layout += struct.pack("<I", a.dtype.itemsize)
        msgs = _message(0x01, space) + _message(0x03, _dtype_message(a.dtype)) + _message(0x08, layout)
        nmsg = 3
        pipeline = []
        if is_chunked and shuffle:
            pipeline.append(struct.pack("<HHHHI4x", 2, 0, 1, 1, a.dtype.itemsize))
        if is_chunked and gzip is not None:
            pipeline.append(struct.pack("<HHHHI4x", 1, 0, 1, 1, gzip))
        if pipeline:
            msgs += _message(0x0B, struct.pack("<BB6x", 1, len(pipeline)) + b"".join(pipeline))
            nmsg += 1
        headers[name] = alloc(struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)

    # root group: local heap, one B-tree leaf, one symbol node
    heap_data = bytearray(8)
    name_off = {}
    for name in names:
        name_off[name] = len(heap_data)
        raw = name.encode("ascii") + b"\0"
        heap_data.extend(raw + b"\0" * (-len(raw) % 8))
    heap_seg = alloc(bytes(heap_data))
    heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_seg))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for name in names:
        snod += struct.pack("<QQII16x", name_off[name], headers[name], 0, 0)
    snod_at = alloc(snod + b"\0" * (8 + 2 * _LEAF_K * 40 - len(snod)))
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF)
    tree += struct.pack("<QQQ", 0, snod_at, name_off[names[-1]] if names else 0)
    tree_at = alloc(tree + b"\0" * (24 + 2 * _INTERNAL_K * 16 + 8 - len(tree)))
    root_msgs = _message(0x11, struct.pack("<QQ", tree_at, heap))
    root_header = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs)
    out.extend(b"\0" * (-len(out) % 8))
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, len(out), _UNDEF)
    sb += struct.pack("<QQII", 0, root_header, 1, 0) + struct.pack("<QQ", tree_at, heap)
    assert len(sb) == 96
    out[0:96] = sb
    with open(path, "wb") as f:
        f.write(out)
