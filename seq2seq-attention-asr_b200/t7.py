"""Torch7 serialisation (`torch.save` / `torch.load`, binary mode) and the reference's checkpoint / log files -- SURVEY 8(f)4.

The training scripts persist their state with `torch.save(paths.concat(opt.savedir,'model.t7'), model)` and
`hdf5.open(...,'log.h5')` (timit/timit.lua:540-562; librispeech/train.lua has the same shape).  `model` is a Lua table
{autoencoder, encoder, decoder, optimConfig, optimState, gradnoise, AWN, opt, ...} (timit/timit.lua:85-95): nn / nngraph
objects whose `.weight` / `.bias` tensors are all views into the ONE flat storage `getParameters()` made
(timit/timit.lua:172), plus plain tables of numbers and tensors.

This module implements the published Torch7 file format (torch/File.lua: a tagged object stream, little endian)

    int32 tag: 0 nil | 1 number (float64) | 2 string (int32 n + bytes) | 3 table | 4 torch object | 5 boolean (int32)
    table   : int32 reference index, then -- first occurrence only -- int32 n and n (key, value) object pairs
    object  : int32 reference index, then -- first occurrence only -- string "V 1", string class name, class payload:
              torch.<T>Tensor  : int32 nDim, int64 size[nDim], int64 stride[nDim], int64 storageOffset (1-based), storage object
              torch.<T>Storage : int64 n, n raw elements
              any other class  : one table object holding its fields (torch.class default write)

so that
  * `load(path)` reads ANY file the reference wrote -- module graphs come back as `Obj(classname, fields)` trees, tensors as
    numpy views that keep their storage identity -- and `flat_parameters(model)` recovers the flat parameter vector (the storage
    every `.weight` views), which is exactly the layout `s2s_param_count` / `s2s_param_segments` use;
  * `save(path, obj)` writes dict / list / number / str / bool / None / numpy arrays (as torch.<T>Tensor over a shared
    torch.<T>Storage when they are views of one base array) / `Obj`, so `save_checkpoint` produces a file `torch.load` opens:
    {parameters, optimConfig, optimState = {paramVariance, paramStd, delta, accDelta} (optim.adadelta's state names), gradnoise,
    AWN, opt, cfg}.  A Lua caller restores with `parameters:copy(ckpt.parameters)` -- the same call timit.lua:229 uses.
Functions (tag 6-8: compiled Lua chunks) are skipped on load and cannot be written.

`write_log` / `read_log` produce / read `log.h5` with the layout of timit/timit.lua:540-551 through the HDF5 subset of h5.py.
"""
import struct

import numpy as np

from . import h5

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_DTYPES = {"Float": np.float32, "Double": np.float64, "Int": np.int32, "Long": np.int64, "Short": np.int16, "Byte": np.uint8,
           "Char": np.int8, "Half": np.float16, "Cuda": np.float32, "CudaDouble": np.float64, "CudaInt": np.int32,
           "CudaLong": np.int64, "CudaByte": np.uint8, "CudaHalf": np.float16}
_NAMES = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int32): "Int", np.dtype(np.int64): "Long",
          np.dtype(np.int16): "Short", np.dtype(np.uint8): "Byte", np.dtype(np.int8): "Char", np.dtype(np.float16): "Half"}


class T7Error(RuntimeError):
    pass


class Obj:
    """a torch.class instance that is not a tensor / storage: class name + field table"""

    def __init__(self, classname, fields=None):
        self.classname = classname
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default) if isinstance(self.fields, dict) else default

    def __repr__(self):
        keys = list(self.fields)[:8] if isinstance(self.fields, dict) else "..."
        return f"Obj({self.classname}, fields={keys})"


class Function:
    """placeholder for a serialised Lua function (not interpreted)"""


def _tensor_kind(classname):
    if not classname.startswith("torch."):
        return None, None
    base = classname[6:]
    for suffix in ("Tensor", "Storage"):
        if base.endswith(suffix) and base[:-len(suffix)] in _DTYPES:
            return suffix, _DTYPES[base[:-len(suffix)]]
    return None, None


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.p = 0
        self.refs = {}

    def _unpack(self, fmt):
        n = struct.calcsize(fmt)
        if self.p + n > len(self.b):
            raise T7Error(f"truncated file at byte {self.p}")
        v = struct.unpack_from(fmt, self.b, self.p)
        self.p += n
        return v[0] if len(v) == 1 else v

    def i32(self):
        return self._unpack("<i")

    def i64(self):
        return self._unpack("<q")

    def string(self):
        n = self.i32()
        if n < 0 or self.p + n > len(self.b):
            raise T7Error(f"bad string length {n} at byte {self.p}")
        s = bytes(self.b[self.p:self.p + n])
        self.p += n
        return s.decode("latin-1")

    def obj(self):
        tag = self.i32()
        if tag == TYPE_NIL:
            return None
        if tag == TYPE_NUMBER:
            v = self._unpack("<d")
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if tag == TYPE_STRING:
            return self.string()
        if tag == TYPE_BOOLEAN:
            return self.i32() != 0
        if tag == TYPE_TABLE:
            idx = self.i32()
            if idx in self.refs:
                return self.refs[idx]
            out = {}
            self.refs[idx] = out
            n = self.i32()
            for _ in range(n):
                k = self.obj()
                v = self.obj()
                if isinstance(k, (dict, list, np.ndarray)):
                    k = id(k)
                out[k] = v
            return out
        if tag == TYPE_TORCH:
            idx = self.i32()
            if idx in self.refs:
                return self.refs[idx]
            version = self.string()
            classname = self.string() if version.startswith("V ") else version
            kind, dtype = _tensor_kind(classname)
            if kind == "Storage":
                n = self.i64()
                nbytes = n * np.dtype(dtype).itemsize
                if n < 0 or self.p + nbytes > len(self.b):
                    raise T7Error(f"{classname}: {n} elements do not fit the file")
                arr = np.frombuffer(self.b, dtype=np.dtype(dtype).newbyteorder("<"), count=n, offset=self.p).astype(dtype)
                self.p += nbytes
                self.refs[idx] = arr
                return arr
            if kind == "Tensor":
                nd = self.i32()
                size = [self.i64() for _ in range(nd)]
                stride = [self.i64() for _ in range(nd)]
                off = self.i64() - 1
                self.refs[idx] = None                 # (a tensor cannot contain itself; keeps the index reserved)
                storage = self.obj()
                if storage is None or nd == 0:
                    t = np.zeros([0] * max(nd, 1), dtype=dtype)
                else:
                    item = storage.dtype.itemsize
                    need = off + sum((n - 1) * st for n, st in zip(size, stride)) + 1 if all(n > 0 for n in size) else 0
                    if off < 0 or need > storage.size or any(st < 0 for st in stride):
                        raise T7Error(f"{classname}: size {size} stride {stride} offset {off + 1} exceeds its storage of {storage.size}")
                    # a view whose .base is the storage array: tensors of one storage stay recognisable as such (flat_parameters)
                    t = np.ndarray(shape=size, dtype=storage.dtype, buffer=storage, offset=off * item, strides=[st * item for st in stride]) \
                        if need else np.zeros(size, dtype=dtype)
                self.refs[idx] = t
                return t
            o = Obj(classname)
            self.refs[idx] = o
            o.fields = self.obj()                    # default torch.class serialisation: the field table
            return o
        if tag in (TYPE_FUNCTION, TYPE_RECUR_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION):
            if tag != TYPE_FUNCTION:
                idx = self.i32()
                if idx in self.refs:
                    return self.refs[idx]
                self.refs[idx] = Function()
            n = self.i32()                            # dumped chunk
            self.p += n
            self.obj()                                # upvalues
            return Function()
        raise T7Error(f"unknown type tag {tag} at byte {self.p - 4}")


def load(path):
    """torch.load(path): tables -> dict (keys 1..n kept as ints: see `as_list`), tensors -> numpy, other objects -> Obj"""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    out = r.obj()
    if r.p != len(buf):
        raise T7Error(f"{len(buf) - r.p} trailing bytes after the root object")
    return out


def as_list(table):
    """a Lua array table {1: a, 2: b, ...} as a Python list"""
    n = len(table)
    if sorted(table) != list(range(1, n + 1)):
        raise T7Error("not an array table")
    return [table[i] for i in range(1, n + 1)]


def _storage_base(a):
    base = a
    while isinstance(base, np.ndarray) and base.base is not None and isinstance(base.base, np.ndarray):
        base = base.base
    return base


class _Writer:
    def __init__(self):
        self.out = bytearray()
        self.ids = {}
        self.next = 1
        self.keep = []

    def i32(self, v):
        self.out += struct.pack("<i", v)

    def i64(self, v):
        self.out += struct.pack("<q", v)

    def string(self, s):
        b = s.encode("latin-1")
        self.i32(len(b))
        self.out += b

    def ref(self, o, kind="obj"):
        """(index, first occurrence?)"""
        key = (kind, id(o))          # an array that owns its data is both a tensor and that tensor's storage
        if key in self.ids:
            return self.ids[key], False
        self.ids[key] = self.next
        self.keep.append(o)
        self.next += 1
        return self.ids[key], True

    def storage(self, base):
        name = _NAMES.get(base.dtype)
        if name is None:
            raise T7Error(f"dtype {base.dtype} has no Torch7 storage type")
        self.i32(TYPE_TORCH)
        idx, new = self.ref(base, "storage")
        self.i32(idx)
        if not new:
            return
        self.string("V 1")
        self.string(f"torch.{name}Storage")
        flat = np.ascontiguousarray(base).reshape(-1)
        self.i64(flat.size)
        self.out += flat.astype(base.dtype.newbyteorder("<"), copy=False).tobytes()

    def obj(self, o):
        if o is None:
            self.i32(TYPE_NIL)
        elif isinstance(o, (bool, np.bool_)):
            self.i32(TYPE_BOOLEAN); self.i32(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.i32(TYPE_NUMBER); self.out += struct.pack("<d", float(o))
        elif isinstance(o, str):
            self.i32(TYPE_STRING); self.string(o)
        elif isinstance(o, np.ndarray):
            name = _NAMES.get(o.dtype)
            if name is None:
                raise T7Error(f"dtype {o.dtype} has no Torch7 tensor type")
            self.i32(TYPE_TORCH)
            idx, new = self.ref(o)
            self.i32(idx)
            if not new:
                return
            self.string("V 1")
            self.string(f"torch.{name}Tensor")
            base = _storage_base(o)
            if not (base.flags.c_contiguous and base.dtype == o.dtype):
                base = np.ascontiguousarray(o); o = base; self.keep.append(base)
            item = o.dtype.itemsize
            off = (o.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // item
            self.i32(o.ndim)
            for s in o.shape:
                self.i64(s)
            for s in o.strides:
                self.i64(s // item)
            self.i64(off + 1)
            self.storage(base)
        elif isinstance(o, (list, tuple)):
            self.obj({i + 1: v for i, v in enumerate(o)})
        elif isinstance(o, dict):
            self.i32(TYPE_TABLE)
            idx, new = self.ref(o)
            self.i32(idx)
            if not new:
                return
            self.i32(len(o))
            for k, v in o.items():
                self.obj(k)
                self.obj(v)
        elif isinstance(o, Obj):
            self.i32(TYPE_TORCH)
            idx, new = self.ref(o)
            self.i32(idx)
            if not new:
                return
            self.string("V 1")
            self.string(o.classname)
            self.obj(o.fields)
        else:
            raise T7Error(f"cannot serialise {type(o).__name__}")


def save(path, obj):
    """torch.save(path, obj) for dict / list / number / str / bool / None / numpy / Obj trees (shared objects written once)"""
    w = _Writer()
    w.obj(obj)
    with open(path, "wb") as f:
        f.write(bytes(w.out))


# ---- the reference's checkpoint -------------------------------------------------------------------------------------------------
def _walk(o, seen, fn):
    if id(o) in seen:
        return
    seen.add(id(o))
    if isinstance(o, Obj):
        fn(o)
        _walk(o.fields, seen, fn)
    elif isinstance(o, dict):
        for v in o.values():
            _walk(v, seen, fn)


def flat_parameters(model):
    """The flat parameter vector of a loaded `model` table / module graph: after `getParameters()` (timit/timit.lua:172) every
    `.weight` / `.bias` of the graph views one storage, which the file holds once -- found as the storage most module tensors share
    -- returned together with {(classname, field): (offset, shape)} of the views for cross-checking a layout."""
    views = []

    def visit(o):
        if isinstance(o.fields, dict):
            for name in ("weight", "bias"):
                t = o.fields.get(name)
                if isinstance(t, np.ndarray) and t.size > 0:
                    views.append((o.classname, name, t))
    _walk(model, set(), visit)
    if not views:
        raise T7Error("no module with a weight / bias tensor in this object")
    count = {}
    for _, _, t in views:
        b = _storage_base(t)
        count[id(b)] = (count.get(id(b), (0, b))[0] + 1, b)
    n, base = max(count.values(), key=lambda cb: cb[0])
    layout = []
    for cls, name, t in views:
        if _storage_base(t) is base:
            off = (t.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // t.dtype.itemsize
            layout.append((off, cls, name, tuple(t.shape)))
    layout.sort()
    return np.array(base, copy=True).reshape(-1), layout


def save_checkpoint(path, parameters, optimConfig=None, optimState=None, gradnoise=None, AWN=None, opt=None, cfg=None, extra=None):
    """model.t7 in the reference's shape (timit/timit.lua:85-95, 552) at the level a Lua caller can consume without this library:
    parameters (flat FloatTensor, `parameters:copy(ckpt.parameters)`), optimConfig / optimState (optim.adadelta's names:
    paramVariance, paramStd, delta, accDelta -- `v`, `a` of s2s_adadelta), gradnoise, AWN, opt, cfg"""
    state = None
    if optimState is not None:
        state = {}
        for k, v in optimState.items():
            state[{"v": "paramVariance", "a": "accDelta"}.get(k, k)] = np.asarray(v, dtype=np.float32) if hasattr(v, "__len__") else v
    model = {"parameters": np.ascontiguousarray(np.asarray(parameters, dtype=np.float32).reshape(-1))}
    for k, v in (("optimConfig", optimConfig), ("optimState", state), ("gradnoise", gradnoise), ("AWN", AWN), ("opt", opt), ("cfg", cfg)):
        if v is not None:
            model[k] = v
    if extra:
        model.update(extra)
    save(path, model)


def load_checkpoint(path):
    """Either a file `save_checkpoint` wrote, or a reference model.t7 (module graphs): returns a dict with at least `parameters`
    (flat float32) and whatever of optimConfig / optimState ({v, a, ...}) / gradnoise / AWN / opt the file holds."""
    m = load(path)
    if not isinstance(m, dict):
        raise T7Error("the root object of a checkpoint is a table")
    out = {k: m[k] for k in ("optimConfig", "gradnoise", "AWN", "opt", "cfg") if k in m}
    if isinstance(m.get("parameters"), np.ndarray):
        out["parameters"] = np.array(m["parameters"], dtype=np.float32).reshape(-1)
    else:
        root = m.get("autoencoder", m)
        out["parameters"], out["layout"] = flat_parameters(root)
        out["parameters"] = out["parameters"].astype(np.float32)
    st = m.get("optimState")
    if isinstance(st, dict):
        out["optimState"] = {{"paramVariance": "v", "accDelta": "a"}.get(k, k): (np.array(v).reshape(-1) if isinstance(v, np.ndarray) else v)
                             for k, v in st.items()}
    return out


# ---- log.h5 (timit/timit.lua:540-551) -------------------------------------------------------------------------------------------
def write_log(path, train, valid, alpha_train=None, alpha_valid=None, Ws_train=None, Ws_valid=None, Vh_train=None, Vh_valid=None,
              output=None):
    """train = {accuracy, nll, gradnorms}, valid = {accuracy, nll, PER}: per-epoch float64 vectors (updateLog / updateList,
    timit.lua:420-445); the attention snapshots are float32 like the reference's `:float()` tensors"""
    tree = {"train": {k: np.asarray(v, dtype=np.float64).reshape(-1) for k, v in train.items()},
            "valid": {k: np.asarray(v, dtype=np.float64).reshape(-1) for k, v in valid.items()}}
    for k, v in (("alpha_train", alpha_train), ("alpha_valid", alpha_valid), ("Ws_train", Ws_train), ("Ws_valid", Ws_valid),
                 ("Vh_train", Vh_train), ("Vh_valid", Vh_valid), ("output", output)):
        if v is not None:
            tree[k] = np.asarray(v, dtype=np.float32)
    h5.write(path, tree)


def read_log(path):
    return h5.read(path)


def update_log(log, accuracy, nll, gradnorms=None):
    """updateLog of timit/timit.lua:428-445: append the epoch's numbers"""
    log = dict(log) if log else {}
    log["accuracy"] = np.concatenate([np.asarray(log.get("accuracy", []), dtype=np.float64), [accuracy]])
    log["nll"] = np.concatenate([np.asarray(log.get("nll", []), dtype=np.float64), [nll]])
    if gradnorms is not None:
        log["gradnorms"] = np.concatenate([np.asarray(log.get("gradnorms", []), dtype=np.float64), np.asarray(gradnorms, dtype=np.float64).reshape(-1)])
    return log
