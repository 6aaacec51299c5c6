"""Minimal HDF5 reader / writer for the reference's data files -- the image has no h5py / libhdf5.

The reference stores its corpora as HDF5 written by h5py with default settings (timit/preprocess_timit.py:341-363,
librispeech/preprocess.py:230-236) and reads them with torch-hdf5 `file:all()` (timit/timit.lua:42-43): a tree of groups whose
leaves are plain numeric arrays

    TIMIT       /{train,valid,test}/<k>/{x [L,123] float64, y [T] int64, y39, start, finish}
    LibriSpeech /<i>/{x, chars, words}

That is the classic on-disk subset of the format, which is what this module understands (HDF5 File Format Specification v1.x/2.x):
superblock v0/v1 (v2/v3 with compact link messages), version-1 object headers with continuation blocks, old-style groups
(symbol-table message -> B-tree v1 + local heap + SNOD nodes), dataspace v1/v2, fixed-point / floating-point / fixed-length string
datatypes, data layout v1-v3: compact, contiguous and chunked (B-tree v1 chunk index, deflate / shuffle / fletcher32 filters).
Anything else (dense groups in fractal heaps, compound / variable-length types, virtual datasets) raises H5Error with the construct
named -- it never guesses.  The reader is pinned on a file produced by the real HDF5 library (tests/test_data_path.py).

`write()` produces the same subset (superblock v0, symbol-table groups, contiguous little-endian datasets) so that tests and
synthetic corpora can be generated here and read back by real HDF5 tools.
"""
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.base = 0
        self.O = 8      # size of offsets
        self.L = 8      # size of lengths

    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def addr(self, off):
        a = self.u(off, self.O)
        return None if a == (1 << (8 * self.O)) - 1 else a + self.base

    # ---- superblock -----------------------------------------------------------------------------------
    def open(self):
        off = 0
        while True:
            if self.b[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.b):
                raise H5Error("not an HDF5 file: no superblock signature")
        ver = self.b[off + 8]
        if ver in (0, 1):
            self.O, self.L = self.b[off + 13], self.b[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            base = self.u(p, self.O)
            self.base = base if base else 0
            p += 4 * self.O                         # base, free-space, end-of-file, driver-info addresses
            # root group symbol table entry
            root_hdr = self.addr(p + self.O)
            return ("v1", root_hdr)
        if ver in (2, 3):
            self.O, self.L = self.b[off + 9], self.b[off + 10]
            p = off + 12
            self.base = self.u(p, self.O)
            root_hdr = self.addr(p + 3 * self.O)
            return ("v2", root_hdr)
        raise H5Error(f"unsupported superblock version {ver}")

    # ---- object headers -------------------------------------------------------------------------------
    def messages(self, hdr):
        """[(type, flags, payload offset, payload size)] of the object header at `hdr` (v1 and v2 headers)"""
        out = []
        if self.b[hdr:hdr + 4] == b"OHDR":
            ver, flags = self.b[hdr + 4], self.b[hdr + 5]
            if ver != 2:
                raise H5Error(f"object header version {ver}")
            p = hdr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            size0 = self.u(p, nsz); p += nsz
            blocks = [(p, size0)]
            track = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = self.b[p], self.u(p + 1, 2), self.b[p + 3]
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:
                        caddr, clen = self.addr(p), self.u(p + self.O, self.L)
                        if self.b[caddr:caddr + 4] != b"OCHK":
                            raise H5Error("bad object header continuation chunk")
                        blocks.append((caddr + 4, clen - 8))
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        ver = self.b[hdr]
        if ver != 1:
            raise H5Error(f"object header version {ver} at {hdr}")
        nmsg = self.u(hdr + 2, 2)
        size = self.u(hdr + 8, 4)
        blocks = [(hdr + 16, size)]
        while blocks and nmsg > 0:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and nmsg > 0:
                mtype, msize, mflags = self.u(p, 2), self.u(p + 2, 2), self.b[p + 4]
                p += 8
                nmsg -= 1
                if mtype == 0x10:
                    blocks.append((self.addr(p), self.u(p + self.O, self.L)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # ---- groups -----------------------------------------------------------------------------------------
    def heap_name(self, heap, off):
        if self.b[heap:heap + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = self.addr(heap + 8 + 2 * self.L)
        end = data + off
        while self.b[end] != 0:
            end += 1
        return bytes(self.b[data + off:end]).decode("utf-8")

    def btree_group(self, node, heap, out):
        if self.b[node:node + 4] != b"TREE":
            raise H5Error("bad group B-tree node")
        level, used = self.b[node + 5], self.u(node + 6, 2)
        p = node + 8 + 2 * self.O
        for i in range(used):
            p += self.L                              # key i
            child = self.addr(p); p += self.O
            if level > 0:
                self.btree_group(child, heap, out)
            else:
                if self.b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol table node")
                n = self.u(child + 6, 2)
                q = child + 8
                for _ in range(n):
                    name = self.heap_name(heap, self.u(q, self.O))
                    out[name] = self.addr(q + self.O)
                    q += 2 * self.O + 24

    def links(self, hdr):
        """{name: object header address} for a group, None if the object is not a group"""
        out, is_group = {}, False
        for mtype, _, p, size in self.messages(hdr):
            if mtype == 0x11:                        # symbol table message
                is_group = True
                self.btree_group(self.addr(p), self.addr(p + self.O), out)
            elif mtype == 0x06:                      # link message (compact new-style group)
                is_group = True
                ver, flags = self.b[p], self.b[p + 1]
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.b[q]; q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                nsz = 1 << (flags & 3)
                nlen = self.u(q, nsz); q += nsz
                name = bytes(self.b[q:q + nlen]).decode("utf-8"); q += nlen
                if ltype != 0:
                    raise H5Error(f"link '{name}': only hard links are supported")
                out[name] = self.addr(q)
            elif mtype == 0x02:                      # link info: dense storage lives in a fractal heap
                is_group = True
                ver, flags = self.b[p], self.b[p + 1]
                q = p + 2 + (8 if flags & 1 else 0)
                if self.addr(q) is not None:
                    raise H5Error("group with dense link storage (fractal heap): not supported; rewrite the file with libver='earliest'")
        return out if is_group else None

    # ---- datasets ---------------------------------------------------------------------------------------
    def dtype_of(self, p):
        cls, ver = self.b[p] & 0x0F, self.b[p] >> 4
        bits = self.u(p + 1, 3)
        size = self.u(p + 4, 4)
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits & 0x08 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Error(f"floating-point datatype of {size} bytes")
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        names = {2: "time", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum", 9: "variable-length", 10: "array"}
        raise H5Error(f"datatype class {cls} ({names.get(cls, '?')}): only integer, float and fixed-length string datasets are supported")

    def dataset(self, hdr):
        shape = dtype = layout = None
        filters = []
        for mtype, _, p, size in self.messages(hdr):
            if mtype == 0x01:
                ver, rank = self.b[p], self.b[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self.u(q + i * self.L, self.L) for i in range(rank))
            elif mtype == 0x03:
                dtype = self.dtype_of(p)
            elif mtype == 0x08:
                ver = self.b[p]
                if ver == 3:
                    cls = self.b[p + 1]
                    if cls == 0:
                        n = self.u(p + 2, 2)
                        layout = ("compact", p + 4, n)
                    elif cls == 1:
                        layout = ("contiguous", self.addr(p + 2), self.u(p + 2 + self.O, self.L))
                    elif cls == 2:
                        nd = self.b[p + 2]
                        bt = self.addr(p + 3)
                        dims = tuple(self.u(p + 3 + self.O + 4 * i, 4) for i in range(nd))
                        layout = ("chunked", bt, dims)
                    else:
                        raise H5Error(f"data layout class {cls}")
                elif ver in (1, 2):
                    nd, cls = self.b[p + 1], self.b[p + 2]
                    q = p + 8
                    a = None
                    if cls != 0:
                        a = self.addr(q); q += self.O
                    dims = tuple(self.u(q + 4 * i, 4) for i in range(nd)); q += 4 * nd
                    if cls == 1:
                        layout = ("contiguous", a, None)
                    elif cls == 2:
                        es = self.u(q, 4)
                        layout = ("chunked", a, dims + (es,))
                    else:
                        n = self.u(q, 4)
                        layout = ("compact", q + 4, n)
                else:
                    raise H5Error(f"data layout message version {ver}")
            elif mtype == 0x0B:
                ver, nf = self.b[p], self.b[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u(q, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self.u(q + 2, 2); q += 4
                    else:
                        nlen = 0; q += 2
                    ncd = self.u(q + 2, 2); q += 4
                    if ver == 1:
                        nlen = (nlen + 7) & ~7
                    q += nlen
                    cd = [self.u(q + 4 * i, 4) for i in range(ncd)]
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = self.b[layout[1]:layout[1] + layout[2]]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        if layout[0] == "contiguous":
            if layout[1] is None:
                return np.zeros(shape, dtype=dtype)         # never written: fill value
            raw = self.b[layout[1]:layout[1] + n * dtype.itemsize]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        # chunked: walk the chunk B-tree
        _, bt, cdims = layout
        cshape = cdims[:-1]
        out = np.zeros(shape, dtype=dtype)
        if bt is not None:
            self.btree_chunks(bt, len(cdims), cshape, dtype, filters, out)
        return out

    def btree_chunks(self, node, nd, cshape, dtype, filters, out):
        if self.b[node:node + 4] != b"TREE":
            raise H5Error("bad chunk B-tree node")
        level, used = self.b[node + 5], self.u(node + 6, 2)
        p = node + 8 + 2 * self.O
        ksz = 8 + 8 * nd
        for i in range(used):
            csize, fmask = self.u(p, 4), self.u(p + 4, 4)
            offs = tuple(self.u(p + 8 + 8 * j, 8) for j in range(nd - 1))
            child = self.addr(p + ksz)
            p += ksz + self.O
            if level > 0:
                self.btree_chunks(child, nd, cshape, dtype, filters, out)
                continue
            raw = bytes(self.b[child:child + csize])
            for k, (fid, cd) in reversed(list(enumerate(filters))):
                if fmask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    nel = len(a) // es
                    raw = a[:nel * es].reshape(es, nel).T.tobytes() + a[nel * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error(f"filter id {fid}: only deflate, shuffle and fletcher32 are supported")
            chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cshape))).reshape(cshape)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs, cshape, out.shape):
                e = min(o + c, s)
                sl_out.append(slice(o, e)); sl_in.append(slice(0, e - o))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]

    def tree(self, hdr, lazy=False):
        kids = self.links(hdr)
        if kids is None:
            return self.dataset(hdr)
        return {k: self.tree(a) for k, a in kids.items()}


def read(path):
    """The whole file as nested dicts of numpy arrays -- what torch-hdf5's `hdf5.open(f):all()` returns (timit/timit.lua:42-43)."""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(memoryview(buf))
    _, root = r.open()
    return r.tree(root)


class File:
    """Lazy view: `f.keys(path)` lists a group, `f[path]` reads one dataset (corpora are read utterance by utterance)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self._buf = f.read()
        self._r = _Reader(memoryview(self._buf))
        _, self._root = self._r.open()
        self._links = {}

    def _resolve(self, path):
        hdr = self._root
        for part in [p for p in path.split("/") if p]:
            if hdr not in self._links:
                self._links[hdr] = self._r.links(hdr)
            kids = self._links[hdr]
            if kids is None or part not in kids:
                raise KeyError(path)
            hdr = kids[part]
        return hdr

    def keys(self, path="/"):
        hdr = self._resolve(path)
        if hdr not in self._links:
            self._links[hdr] = self._r.links(hdr)
        if self._links[hdr] is None:
            raise KeyError(f"{path} is a dataset")
        return list(self._links[hdr].keys())

    def __getitem__(self, path):
        hdr = self._resolve(path)
        kids = self._r.links(hdr)
        return self._r.dataset(hdr) if kids is None else self._r.tree(hdr)


# ---- writer (superblock v0, symbol-table groups, contiguous little-endian datasets: the h5py-default subset) ---------------------
class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b"\0" * pad
        off = len(self.buf)
        self.buf += b"\0" * n
        return off

    def put(self, off, data):
        self.buf[off:off + len(data)] = data

    @staticmethod
    def msg(mtype, payload, flags=0):
        pad = (-len(payload)) % 8
        return struct.pack("<HHB3x", mtype, len(payload) + pad, flags) + payload + b"\0" * pad

    def header(self, msgs):
        body = b"".join(msgs)
        off = self.alloc(16 + len(body))
        self.put(off, struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)
        return off

    def dataset(self, arr):
        arr = np.asarray(arr)
        arr = arr if arr.ndim == 0 else np.ascontiguousarray(arr)     # (ascontiguousarray would promote a scalar to 1-d)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        k = arr.dtype.kind
        if k in "iu":
            dt = struct.pack("<B3BI", 0x10 | 0, 0x08 if k == "i" else 0x00, 0, 0, arr.itemsize) + struct.pack("<HH", 0, 8 * arr.itemsize)
        elif k == "f" and arr.itemsize in (4, 8):
            # IEEE little-endian: byte order 0, padding 0, mantissa normalisation 2 (implied msb), sign location in byte 1
            sign = 8 * arr.itemsize - 1
            if arr.itemsize == 4:
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            else:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            dt = struct.pack("<B3BI", 0x10 | 1, 0x20, sign, 0, arr.itemsize) + props
        else:
            raise H5Error(f"write: dtype {arr.dtype} not supported")
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        data = arr.tobytes()
        daddr = self.alloc(len(data)) if data else UNDEF
        if data:
            self.put(daddr, data)
        layout = struct.pack("<BBQQ", 3, 1, daddr, len(data))
        fill = struct.pack("<BBBB", 2, 2, 0, 0)              # fill value message v2: allocate late, never write, undefined
        return self.header([self.msg(0x01, space), self.msg(0x03, dt, 1), self.msg(0x05, fill), self.msg(0x08, layout)])

    def group(self, entries):
        """entries: {name: object header address}; returns (header address, btree address, heap address)"""
        names = sorted(entries)
        heap_data = bytearray(b"\0" * 8)                     # offset 0 = empty string (the B-tree's first key)
        noff = {}
        for n in names:
            noff[n] = len(heap_data)
            heap_data += n.encode("utf-8") + b"\0"
            heap_data += b"\0" * ((-len(heap_data)) % 8)
        hd = self.alloc(len(heap_data))
        self.put(hd, bytes(heap_data))
        heap = self.alloc(32)
        self.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
        # symbol table nodes of up to 2*K entries (K = the superblock's group leaf node K), one B-tree node above them
        K = self.leaf_k
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 2 * K):
            chunk = names[i:i + 2 * K]
            s = self.alloc(8 + 2 * K * 40)
            body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", noff[n], entries[n], 0, 0)
            self.put(s, body)
            snods.append(s)
            keys.append(noff[chunk[-1]] if chunk else 0)
        if len(snods) > 2 * self.internal_k:
            raise H5Error("write: group too large for one B-tree node; raise internal_k")
        bt = self.alloc(24 + (2 * self.internal_k + 1) * 8 + 2 * self.internal_k * 8)
        body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, s in enumerate(snods):
            body += struct.pack("<QQ", keys[i], s)
        body += struct.pack("<Q", keys[len(snods)])
        self.put(bt, body)
        hdr = self.header([self.msg(0x11, struct.pack("<QQ", bt, heap))])
        return hdr, bt, heap

    def node(self, obj):
        if isinstance(obj, dict):
            hdr, _, _ = self.group({k: self.node(v) for k, v in obj.items()})
            return hdr
        return self.dataset(np.asarray(obj))


def write(path, tree, leaf_k=4, internal_k=16):
    """Write nested dicts of numpy arrays as an HDF5 file in the classic layout h5py / libhdf5 produce by default."""
    largest = [1]

    def walk(t):
        if isinstance(t, dict):
            largest[0] = max(largest[0], len(t))
            for v in t.values():
                walk(v)
    walk(tree)
    w = _Writer()
    w.leaf_k = max(leaf_k, -(-largest[0] // (2 * internal_k * 2)))     # keep every group within one B-tree node
    w.internal_k = internal_k
    w.alloc(96)                                              # superblock v0 (8-byte offsets / lengths) + root symbol table entry
    entries = {k: w.node(v) for k, v in tree.items()}
    root, bt, heap = w.group(entries)
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, w.leaf_k, w.internal_k, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, heap)
    w.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
