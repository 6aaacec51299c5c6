"""ctypes binding of libs2s_b200.so (the C ABI declared in include/s2s_b200.h).

PyTorch is used only as the owner of device memory / streams; every compute call goes through the
C ABI with raw device pointers.  There is no CPU fallback: if the shared library is missing this
module raises, and on a machine without a CUDA device `Context()` raises.
"""
import ctypes as C
import os
import re

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libs2s_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_DIR), "include", "s2s_b200.h")

CFG_FIELDS = ("D", "H", "NL", "S", "ST", "V", "K", "KF", "M", "MW", "MLP")
CFG_DEFAULTS = {"MLP": 1}


class VggCfg(C.Structure):
    """s2s_vgg_cfg: planes of conv1-2 / conv3-4, width of the 1x1 stack, annotation depth (model_vgg.lua:24-54)"""
    _fields_ = [(k, C.c_int) for k in ("C1", "C2", "HID", "OUT")]

    @classmethod
    def of(cls, cfg):
        return cfg if isinstance(cfg, cls) else cls(*[int(cfg[k]) for k in ("C1", "C2", "HID", "OUT")])


class ModelCfg(C.Structure):
    _fields_ = [(k, C.c_int) for k in CFG_FIELDS]

    @classmethod
    def from_dict(cls, d):
        return cls(*[int(d.get(k, CFG_DEFAULTS.get(k))) for k in CFG_FIELDS])


class S2SError(RuntimeError):
    pass


_lib = None


def declared_symbols():
    """Names of every function declared in include/s2s_b200.h"""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(s2s_[a-z0-9_]+)\s*\(", txt)))


def load():
    """Load the shared library (build it with __graft_entry__.build() or `make -C csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S2SError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64
    cfgp = C.POINTER(ModelCfg)
    sig = {
        "s2s_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
        "s2s_ctx_destroy": (i32, [vp]),
        "s2s_ctx_set_stream": (i32, [vp, vp]),
        "s2s_ctx_synchronize": (i32, [vp]),
        "s2s_last_error": (C.c_char_p, []),
        "s2s_version": (i32, []),
        "s2s_ctx_launch_count": (i64, [vp]),
        "s2s_ctx_kernel_count": (i64, [vp, i32]),
        "s2s_ctx_set_graphs": (i32, [vp, i32]),
        "s2s_ctx_profile": (i32, [vp, i32]),
        "s2s_ctx_profile_read": (i32, [vp, C.POINTER(f64), C.POINTER(i64), C.POINTER(f64)]),
        "s2s_orthogonalize": (i32, [vp, vp, i64, i64, vp]),
        "s2s_dp_available": (i32, []),
        "s2s_dp_unique_id": (i32, [vp]),
        "s2s_dp_init": (i32, [vp, i32, i32, vp]),
        "s2s_dp_rank": (i32, [vp]),
        "s2s_dp_world": (i32, [vp]),
        "s2s_dp_allreduce": (i32, [vp, vp, i64]),
        "s2s_dp_broadcast": (i32, [vp, vp, i64, i32]),
        "s2s_dp_set_overlap": (i32, [vp, i32]),
        "s2s_dp_destroy": (i32, [vp]),
        "s2s_param_count": (i64, [cfgp]),
        "s2s_param_segments": (i32, [cfgp, vp, i32]),
        "s2s_decoder_param_offset": (i64, [cfgp]),
        "s2s_tconv_zb_forward": (i32, [vp, vp, i64, i32, vp, i32, vp]),
        "s2s_tconv_zb_backward": (i32, [vp, vp, i64, i32, vp, i32, vp, vp, vp, f32]),
        "s2s_linear_zb_forward": (i32, [vp, vp, i64, i32, vp, i32, vp]),
        "s2s_linear_zb_backward": (i32, [vp, vp, i64, i32, vp, i32, vp, vp, vp, f32]),
        "s2s_gru_seq_save_floats": (i64, [i32, i32, i32, i32]),
        "s2s_gru_seq_forward": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp]),
        "s2s_gru_seq_backward": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp]),
        "s2s_gru_step_forward": (i32, [vp, vp, i32, i32, vp, vp, i32, vp, vp]),
        "s2s_gru_step_backward": (i32, [vp, vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
        "s2s_dropout_mask": (i32, [vp, f32, u64, i64, vp]),
        "s2s_lstm_step_forward": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
        "s2s_lstm_step_backward": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "s2s_lstm_param_count": (i64, [i32, i32, i32]),
        "s2s_lstm_seq_save_floats": (i64, [i32, i32, i32]),
        "s2s_lstm_seq_forward": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp]),
        "s2s_lstm_seq_backward": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp]),
        "s2s_attention_forward": (i32, [vp, cfgp, vp, vp, vp, i32, i32, vp, vp, i32, vp, f32, vp]),
        "s2s_attention_backward": (i32, [vp, cfgp, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, f32, vp, vp]),
        "s2s_attention_get": (i32, [vp, i32, vp]),
        "s2s_attention_step": (i32, [vp, cfgp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]),
        "s2s_beam_search": (i32, [vp, cfgp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "s2s_model_forward": (i32, [vp, cfgp, vp, vp, vp, i32, i32, vp, vp, i32, vp, f32, i32, vp, vp]),
        "s2s_model_fwdbwd": (i32, [vp, cfgp, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, f32, i32, vp, vp, vp]),
        "s2s_model_get_annotations": (i32, [vp, vp]),
        "s2s_weightnoise_sample": (i32, [vp, vp, vp, u64, f32, i64, vp]),
        "s2s_awn_sample": (i32, [vp, vp, vp, u64, i64, vp]),
        "s2s_awn_forward": (i32, [vp, vp, i64, f64, f64, C.POINTER(f64)]),
        "s2s_awn_accgrad": (i32, [vp, vp, vp, i64, f64, vp]),
        "s2s_grad_finalize": (i32, [vp, vp, vp, i64, i32, f64, f64, vp, u64, f64, C.POINTER(f64)]),
        "s2s_adadelta": (i32, [vp, vp, vp, vp, vp, i64, f64, f64]),
        "s2s_rownorm_constraint": (i32, [vp, vp, i64, i64, f64, C.POINTER(i32)]),
        "s2s_model_rownorm_constraint": (i32, [vp, cfgp, vp, f64, C.POINTER(i32)]),
        "s2s_gemm_f32": (i32, [vp, i32, i32, i32, i32, i32, i32, f32, vp, i32, vp, i32, f32, vp, i32, vp]),
        "s2s_attn_step_forward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
        "s2s_attn_step_backward": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "s2s_vgg_param_count": (i64, [vp, i32]),
        "s2s_vgg_out_len": (i32, [i32]),
        "s2s_vgg_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
        "s2s_vgg_backward": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
        "s2s_graph_begin": (i32, [vp]),
        "s2s_graph_end": (i32, [vp, vp]),
        "s2s_graph_launch": (i32, [vp, i32]),
        "s2s_graph_destroy": (i32, [vp, i32]),
        "s2s_conv3_forward": (i32, [vp, vp, i64, i32, i32, vp, vp, i32, vp, i32]),
        "s2s_conv3_dgrad": (i32, [vp, vp, i64, i32, i32, vp, i32, vp]),
        "s2s_conv3_wgrad": (i32, [vp, vp, vp, i64, i32, i32, i32, vp]),
        "s2s_labels_from_onehot": (i32, [vp, vp, i64, i32, vp]),
        "s2s_onehot": (i32, [vp, vp, i64, i32, vp]),
        "s2s_nll_grad_seed": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
        "s2s_edit_distance": (i32, [vp, i32, vp, i32, vp]),
        "s2s_attn_step_forward_loc": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "s2s_attn_step_backward_loc": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    lib._sig = sig
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise S2SError(load().s2s_last_error().decode())
