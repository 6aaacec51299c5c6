"""Torch7 .t7 serialisation and the reference's checkpoint / log files (timit/timit.lua:85-95, 540-562).
The byte stream of `test_reads_a_stream_laid_out_like_torch_save` is assembled by hand from the published format
(torch/File.lua), independently of the writer under test: a model table with two modules whose weights view ONE
FloatStorage (what getParameters() leaves behind), a table referenced twice, a boolean, and an optimState."""
import struct

import numpy as np
import pytest

import s2s_b200 as s2s
from s2s_b200 import t7


def _i32(v): return struct.pack("<i", v)
def _i64(v): return struct.pack("<q", v)
def _str(s): return _i32(len(s)) + s.encode()
def _num(v): return _i32(1) + struct.pack("<d", v)
def _string(s): return _i32(2) + _str(s)


def _tensor(idx, cls, size, stride, off, storage_bytes):
    out = _i32(4) + _i32(idx) + _str("V 1") + _str(cls) + _i32(len(size))
    for s in size: out += _i64(s)
    for s in stride: out += _i64(s)
    return out + _i64(off) + storage_bytes


def test_reads_a_stream_laid_out_like_torch_save(tmp_path):
    flat = np.arange(10, dtype=np.float32) * 0.5
    storage_first = _i32(4) + _i32(100) + _str("V 1") + _str("torch.FloatStorage") + _i64(10) + flat.tobytes()
    storage_again = _i32(4) + _i32(100)                                    # back-reference: index only
    lin = _i32(4) + _i32(2) + _str("V 1") + _str("nn.LinearZeroBias") + _i32(3) + _i32(3) + _i32(2) \
        + _string("weight") + _tensor(4, "torch.FloatTensor", [2, 3], [3, 1], 1, storage_first) \
        + _string("train") + _i32(5) + _i32(1)
    conv = _i32(4) + _i32(5) + _str("V 1") + _str("nn.TemporalConvolutionZeroBias") + _i32(3) + _i32(6) + _i32(2) \
        + _string("weight") + _tensor(7, "torch.FloatTensor", [2, 2], [2, 1], 7, storage_again) \
        + _string("bias") + _i32(0)
    opt_table = _i32(3) + _i32(8) + _i32(1) + _string("rho") + _num(0.95)
    root = _i32(3) + _i32(1) + _i32(4) \
        + _string("decoder") + lin + _string("Vh") + conv \
        + _string("optimConfig") + opt_table + _string("optimConfigAgain") + _i32(3) + _i32(8)
    p = tmp_path / "model.t7"
    p.write_bytes(root)
    m = t7.load(str(p))
    assert m["decoder"].classname == "nn.LinearZeroBias" and m["decoder"]["train"] is True
    np.testing.assert_array_equal(m["decoder"]["weight"], flat[:6].reshape(2, 3))
    np.testing.assert_array_equal(m["Vh"]["weight"], flat[6:10].reshape(2, 2))
    assert m["Vh"]["bias"] is None
    assert m["optimConfig"] is m["optimConfigAgain"] and m["optimConfig"]["rho"] == 0.95
    params, layout = t7.flat_parameters(m)
    np.testing.assert_array_equal(params, flat)
    assert layout == [(0, "nn.LinearZeroBias", "weight", (2, 3)), (6, "nn.TemporalConvolutionZeroBias", "weight", (2, 2))]
    ck = t7.load_checkpoint(str(p))
    np.testing.assert_array_equal(ck["parameters"], flat)


def test_round_trip_keeps_values_types_and_sharing(tmp_path):
    base = np.arange(24, dtype=np.float32)
    shared = {"eps": 1e-8}
    obj = {"w1": base[:12].reshape(3, 4), "w2": base[12:].reshape(2, 6), "d": np.linspace(0, 1, 5), "i": np.arange(4, dtype=np.int64),
           "col": base.reshape(4, 6)[:, 2], "name": "chorowski", "flag": False, "n": 3, "nil": None, "list": [1.5, "a", [2, 3]],
           "mod": t7.Obj("nn.GRU", {"weight": base[:6].reshape(2, 3), "cfg": shared}), "cfg": shared}
    p = str(tmp_path / "x.t7")
    t7.save(p, obj)
    back = t7.load(p)
    for k in ("w1", "w2", "d", "i", "col"):
        np.testing.assert_array_equal(back[k], obj[k])
        assert back[k].dtype == obj[k].dtype
    assert back["name"] == "chorowski" and back["flag"] is False and back["n"] == 3 and back["nil"] is None
    assert t7.as_list(back["list"])[:2] == [1.5, "a"] and t7.as_list(back["list"][3]) == [2, 3]
    assert back["mod"].classname == "nn.GRU" and back["mod"]["cfg"] is back["cfg"]
    # the three views of `base` share ONE storage in the file, as tensors made by getParameters() do
    flat, layout = t7.flat_parameters({"mod": back["mod"]})
    assert flat.size == 24
    raw = open(p, "rb").read()
    assert raw.count(b"torch.FloatStorage") == 1


def test_checkpoint_round_trip_with_adadelta_state(tmp_path):
    cfg = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=16, KF=10, M=64, MW=7)
    P = s2s.init_params(cfg, seed=5)
    v, a = np.abs(P) * 0.1, np.abs(P) * 0.01
    p = str(tmp_path / "model.t7")
    t7.save_checkpoint(p, P, optimConfig={"rho": 0.95, "eps": 1e-8}, optimState={"v": v, "a": a}, gradnoise={"eta": 0.01, "gamma": 0.55, "t": 12},
                       opt={"batchSize": 32, "penalty": 0.0}, cfg=cfg)
    ck = t7.load_checkpoint(p)
    np.testing.assert_array_equal(ck["parameters"], P.astype(np.float32))
    np.testing.assert_array_equal(ck["optimState"]["v"], v.astype(np.float32))
    np.testing.assert_array_equal(ck["optimState"]["a"], a.astype(np.float32))
    assert ck["optimConfig"] == {"rho": 0.95, "eps": 1e-8} and ck["gradnoise"]["t"] == 12 and ck["cfg"]["KF"] == 10
    raw = t7.load(p)                      # the names a Lua optim.adadelta state uses
    assert set(raw["optimState"]) == {"paramVariance", "accDelta"}


def test_truncated_and_foreign_files_fail_loudly(tmp_path):
    p = tmp_path / "bad.t7"
    p.write_bytes(_i32(3) + _i32(1) + _i32(2) + _string("a"))
    with pytest.raises(t7.T7Error):
        t7.load(str(p))
    p.write_bytes(_i32(42))
    with pytest.raises(t7.T7Error):
        t7.load(str(p))


def test_log_h5_layout(tmp_path):
    train = t7.update_log(t7.update_log(None, 0.31, 2.5, gradnorms=[3.0, 2.0]), 0.42, 1.9, gradnorms=[1.5])
    valid = t7.update_log(t7.update_log(None, 0.30, 2.6), 0.40, 2.0)
    valid["PER"] = np.array([0.55, 0.41])
    alpha = np.random.default_rng(0).random((3, 5, 7)).astype(np.float32)
    p = str(tmp_path / "log.h5")
    t7.write_log(p, train, valid, alpha_train=alpha, Ws_valid=alpha[0])
    back = t7.read_log(p)
    np.testing.assert_allclose(back["train"]["accuracy"], [0.31, 0.42])
    np.testing.assert_allclose(back["train"]["gradnorms"], [3.0, 2.0, 1.5])
    np.testing.assert_allclose(back["valid"]["PER"], [0.55, 0.41])
    assert back["train"]["nll"].dtype == np.float64 and back["alpha_train"].dtype == np.float32
    np.testing.assert_array_equal(back["alpha_train"], alpha)
    np.testing.assert_array_equal(back["Ws_valid"], alpha[0])
