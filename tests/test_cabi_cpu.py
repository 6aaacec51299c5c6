"""CPU-side checks of the boundary: the shared library loads, exports every symbol declared in
include/s2s_b200.h, agrees with the oracle on the flat parameter layout, and fails loudly (never falls
back) when there is no CUDA device.  No compute call is made here."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(s2s):
    lib = s2s.load()
    names = s2s.declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.s2s_version() == 100


def test_layout_matches_oracle(s2s, orc64):
    from oracle.oracle import CHOROWSKI_TIMIT, segment_names
    for cfg in (CHOROWSKI_TIMIT, dict(CHOROWSKI_TIMIT, K=16, KF=10), dict(D=13, H=128, NL=2, S=128, ST=64, V=11, K=3, KF=4, M=8, MW=3)):
        assert s2s.param_count(cfg) == orc64.param_count(cfg)
        assert s2s.decoder_param_offset(cfg) > 0
        segs = s2s.param_segments(cfg)
        assert [tuple(int(v) for v in r) for r in orc64.param_segments(cfg)] == [tuple(r) for r in segs]
        assert s2s.segment_names(cfg) == segment_names(cfg)
    assert s2s.param_count(CHOROWSKI_TIMIT) == 4356735          # SURVEY App. A parameter census
    assert s2s.param_count(dict(CHOROWSKI_TIMIT, K=16, KF=10)) == 4365615


def test_invalid_cfg_is_an_error(s2s):
    assert s2s.param_count(dict(D=0, H=1, NL=1, S=1, ST=1, V=2, K=0, KF=1, M=1, MW=1)) == -1
    assert b"invalid model cfg" in s2s.load().s2s_last_error()


def test_no_cpu_fallback(s2s):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(s2s.S2SError, match="no CUDA device"):
        s2s.Context(0)
    h = C.c_void_p()
    assert s2s.load().s2s_ctx_create(0, None, C.byref(h)) != 0
    assert b"no CPU fallback" in s2s.load().s2s_last_error()


def test_product_package_does_not_import_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seq2seq-attention-asr_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".lua")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle's", "").lower() or f in ("_never_",), f"{f} mentions the oracle"


def test_init_params_rules(s2s):
    cfg = dict(D=13, H=128, NL=1, S=128, ST=64, V=11, K=2, KF=4, M=8, MW=3)
    P = s2s.init_params(cfg, seed=3)
    for (off, rows, cols), name in zip(s2s.param_segments(cfg), s2s.segment_names(cfg)):
        seg = P[off:off + rows * cols]
        if name in ("bV", "bU", "be"):
            assert not seg.any()                                  # dead biases stay zero
        elif cols > 1:
            fan = cfg["KF"] if name == "WF" else cols
            assert np.abs(seg).max() <= 1 / np.sqrt(fan) + 1e-7 and seg.std() > 0


def test_edit_distance_matches_wagner_fischer():
    # utils.lua:3-27 restated in Python; the C entry is host-only, so it runs without a GPU
    import numpy as np
    import s2s_b200 as s2s

    def wf(a, b):
        m, n = len(a) + 1, len(b) + 1
        d = np.zeros((m, n), np.int64)
        d[:, 0] = np.arange(m); d[0, :] = np.arange(n)
        for j in range(1, n):
            for i in range(1, m):
                d[i, j] = d[i - 1, j - 1] if a[i - 1] == b[j - 1] else 1 + min(d[i - 1, j], d[i, j - 1], d[i - 1, j - 1])
        return int(d[m - 1, n - 1])

    rng = np.random.default_rng(0)
    assert s2s.edit_distance([], []) == 0 and s2s.edit_distance([1, 2, 3], []) == 3 and s2s.edit_distance([], [4]) == 1
    assert s2s.edit_distance([1, 2, 3], [1, 2, 3]) == 0 and s2s.edit_distance([1, 2, 3], [1, 3]) == 1
    for _ in range(50):
        a = rng.integers(0, 5, rng.integers(0, 30)); b = rng.integers(0, 5, rng.integers(0, 30))
        assert s2s.edit_distance(a, b) == wf(list(a), list(b))


def test_lua_ffi_cdef_matches_the_header():
    """lua/s2s_ffi.lua's cdef block is generated from include/s2s_b200.h: every declared function appears there with the same
    number of parameters (the shims cannot be executed in this image -- no LuaJIT / Torch7 -- so drift is caught textually)."""
    import re
    import s2s_b200 as s2s
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "s2s_b200.h")).read(), flags=re.S)
    lua = open(os.path.join(root, "seq2seq-attention-asr_b200", "lua", "s2s_ffi.lua")).read()
    cdef = lua[lua.index("ffi.cdef[["):lua.index("]]")]

    def protos(txt):
        out = {}
        for m in re.finditer(r"\b(s2s_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
            args = m.group(2).strip()
            out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
        return out
    h, l = protos(hdr), protos(cdef)
    assert set(s2s.declared_symbols()) <= set(l), sorted(set(s2s.declared_symbols()) - set(l))
    assert all(h[k] == l[k] for k in h), {k: (h[k], l[k]) for k in h if h[k] != l.get(k)}


def test_lua_shims_cover_the_reference_module_surface():
    """every custom class of the reference's root library that is on the path has a shim that registers the same torch
    class name, and lua/TrainUtils.lua exports the reference's table (TrainUtils.lua:202-213)"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lua = os.path.join(root, "seq2seq-attention-asr_b200", "lua")
    want = {"Attention": "nn.Attention", "RNNAttention": "nn.RNNAttention", "Recurrent": "nn.Recurrent", "GRU": "nn.GRU", "LSTM": "nn.LSTM",
            "RNN": "nn.RNN", "LinearZeroBias": "nn.LinearZeroBias", "TemporalConvolutionZeroBias": "nn.TemporalConvolutionZeroBias",
            "WeightNoise": "nn.WeightNoise", "AdaptiveWeightNoise": "nn.AdaptiveWeightNoise"}
    for f, cls in want.items():
        txt = open(os.path.join(lua, f + ".lua")).read()
        assert re.search(r"torch\.class\('%s'" % re.escape(cls), txt), f
        for method in ("updateOutput", "updateGradInput"):
            if f not in ("WeightNoise", "AdaptiveWeightNoise"):
                assert ":" + method in txt, (f, method)
    tu = open(os.path.join(lua, "TrainUtils.lua")).read()
    for name in ("orthogonalize", "orthogonalizeGraph", "checkOrthogonalization", "columnNormConstraint", "columnNormConstraintGraph",
                 "checkColumnNormConstraint", "checkColumnNormConstraintGraph", "apply2graph", "getnorms", "checkoutput"):
        assert re.search(r"function T\.%s\b" % name, tu), name
