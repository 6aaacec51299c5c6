"""Data path: the HDF5 subset reader is pinned on a file written by the REAL HDF5 library (a MATLAB v7.3 file that
ships with scipy: 512-byte user block, superblock v0, symbol-table group, version-1 object header), the writer is checked by
round trips, and the bucketed batching reproduces every utterance exactly once per epoch with little padding."""
import importlib
import os

import numpy as np
import pytest

h5 = importlib.import_module("seq2seq-attention-asr_b200.h5")
data = importlib.import_module("seq2seq-attention-asr_b200.data")


def test_reader_on_a_file_written_by_libhdf5():
    scipy_io = pytest.importorskip("scipy.io")
    path = os.path.join(os.path.dirname(scipy_io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's HDF5 test file is not installed")
    t = h5.read(path)
    assert list(t) == ["testdouble"]
    v = t["testdouble"]
    assert v.dtype == np.float64 and v.shape == (9, 1)
    assert np.allclose(v.ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)      # scipy's own expectation for this file: 0:pi/4:2*pi


def test_round_trip_of_the_reference_layout(tmp_path):
    rng = np.random.default_rng(0)
    splits = {"train": [(int(l), int(t)) for l, t in zip(rng.integers(60, 400, 37), rng.integers(5, 60, 37))], "valid": [(80, 9), (121, 30)]}
    path = str(tmp_path / "timit_like.h5")
    tree = data.write_timit_like(path, splits, seed=3)
    back = h5.read(path)
    assert sorted(back) == ["train", "valid"] and len(back["train"]) == 37
    for split in tree:
        for k, u in tree[split].items():
            for name, a in u.items():
                b = back[split][k][name]
                assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b), (split, k, name)
    f = h5.File(path)                                       # lazy access reads single datasets
    assert sorted(f.keys("/train"), key=int) == [str(i) for i in range(37)]
    assert np.array_equal(f["/valid/1/y"], tree["valid"]["1"]["y"])
    # other dtypes / ranks / an empty group / many entries in one group (several symbol-table nodes)
    t2 = {"a": {"f32": rng.standard_normal((3, 4, 5)).astype(np.float32), "u8": np.arange(7, dtype=np.uint8), "i32": np.array([-5, 6], np.int32)},
          "empty": {}, "scalar": np.array(3.5), "big": {str(i): np.array([i], np.int64) for i in range(300)}}
    p2 = str(tmp_path / "misc.h5")
    h5.write(p2, t2)
    b2 = h5.read(p2)
    assert b2["empty"] == {} and b2["scalar"].shape == () and b2["scalar"].item() == 3.5 and len(b2["big"]) == 300 and int(b2["big"]["217"][0]) == 217
    for k, a in t2["a"].items():
        assert b2["a"][k].dtype == a.dtype and np.array_equal(b2["a"][k], a)
    with pytest.raises(h5.H5Error):
        open(str(tmp_path / "junk.h5"), "wb").write(b"not hdf5" * 100)
        h5.read(str(tmp_path / "junk.h5"))


def test_chunked_deflate_datasets_are_read(tmp_path):
    """chunked + deflate + shuffle storage (what h5py writes with compression='gzip', shuffle=True): built by hand against the format
    specification (data layout v3 class 2, filter pipeline v1, B-tree v1 type 1) on top of the writer's primitives"""
    import struct, zlib
    arr = np.arange(7 * 10, dtype=np.float32).reshape(7, 10) * 0.5
    w = h5._Writer(); w.leaf_k, w.internal_k = 4, 16
    w.alloc(96)
    cshape = (4, 8)
    chunks = []
    for r0 in range(0, 7, 4):
        for c0 in range(0, 10, 8):
            c = np.zeros(cshape, np.float32)
            blk = arr[r0:r0 + 4, c0:c0 + 8]
            c[:blk.shape[0], :blk.shape[1]] = blk
            raw = np.frombuffer(c.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()            # shuffle
            z = zlib.compress(raw)                                                             # deflate
            a = w.alloc(len(z)); w.put(a, z)
            chunks.append((len(z), (r0, c0, 0), a))
    bt = w.alloc(24 + 64 * (8 + 8 * 3 + 8))
    body = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(chunks), h5.UNDEF, h5.UNDEF)
    for size, offs, a in chunks:
        body += struct.pack("<II3Q", size, 0, *offs) + struct.pack("<Q", a)
    body += struct.pack("<II3Q", 0, 0, 8, 16, 0)
    w.put(bt, body)
    space = struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", 7, 10)
    dt = struct.pack("<B3BI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 4, 8, 4)
    filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I4x", 4) + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I4x", 6)
    hdr = w.header([w.msg(0x01, space), w.msg(0x03, dt, 1), w.msg(0x0B, filt), w.msg(0x08, layout)])
    root, gbt, heap = w.group({"z": hdr})
    sb = h5.SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<QQQQ", 0, h5.UNDEF, len(w.buf), h5.UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", gbt, heap)
    w.put(0, sb)
    p = str(tmp_path / "chunked.h5")
    open(p, "wb").write(bytes(w.buf))
    assert np.array_equal(h5.read(p)["z"], arr)


def test_bucketed_batches_cover_the_corpus_with_little_padding(tmp_path):
    rng = np.random.default_rng(1)
    utts = [(int(l), int(t)) for l, t in zip(rng.integers(60, 600, 203), rng.integers(5, 70, 203))]
    path = str(tmp_path / "c.h5")
    tree = data.write_timit_like(path, {"train": utts}, seed=5)
    ds = data.TimitH5(path, "train")
    assert len(ds) == 203
    x7, y7 = ds[7]
    assert x7.dtype == np.float32 and np.allclose(x7, tree["train"]["7"]["x"].astype(np.float32)) and np.array_equal(y7, tree["train"]["7"]["y"])
    bb = data.BucketedBatches(ds, 32, shuffle=True, seed=11)
    seen = []
    for batch in bb:
        B = len(batch.index)
        assert batch.X.shape[0] == B and batch.X.shape[1] == batch.lengths.max() and batch.labels.shape[1] == batch.tlens.max()
        for b, j in enumerate(batch.index):
            L, T = utts[j]
            assert batch.lengths[b] == L and batch.tlens[b] == T
            assert np.array_equal(batch.labels[b, :T], tree["train"][str(j)]["y"]) and batch.labels[b, T - 1] == 61      # EOS last
            assert not batch.X[b, L:].any()                                                                               # zero padding
        seen += list(batch.index)
    assert sorted(seen) == list(range(203))                 # every utterance exactly once per epoch
    assert bb.padding_fraction() < 0.12                     # 7 batches over a 60..600 spread; random composition pads > 30% (below)
    rnd = [np.arange(i, min(i + 32, 203)) for i in range(0, 203, 32)]
    lens = ds.lengths()[:, 0]
    assert 1 - sum(lens[b].sum() for b in rnd) / sum(len(b) * lens[b].max() for b in rnd) > 0.3
    # a different seed shuffles the order of the batches, not their composition; two data-parallel ranks split every batch
    o1 = [tuple(b.index) for b in data.BucketedBatches(ds, 32, seed=1)]; o2 = [tuple(b.index) for b in data.BucketedBatches(ds, 32, seed=2)]
    assert o1 != o2 and sorted(o1) == sorted(o2)
    r0 = [b.index for b in data.BucketedBatches(ds, 32, seed=1, world=2, rank=0)]; r1 = [b.index for b in data.BucketedBatches(ds, 32, seed=1, world=2, rank=1)]
    for a, b, full in zip(r0, r1, o1):
        assert tuple(np.concatenate([a, b])) == full and abs(len(a) - len(b)) <= 1
