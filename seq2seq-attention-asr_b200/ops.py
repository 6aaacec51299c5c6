"""Functional host API over the C ABI: every function takes torch CUDA tensors (fp32 / int32,
contiguous), passes their raw device pointers to libs2s_b200.so and returns torch tensors.

Nothing here computes: it is argument marshalling only.  The reference-facing module surface
(nn.Attention, nn.RNN, ...) is in nn.py and is built on these calls.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ModelCfg, S2SError, check

# timit/model_chorowski_baseline.lua:14-46 defaults
CHOROWSKI_TIMIT = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=0, KF=10, M=64, MW=7)

NORMALIZE_NLL = 1
NORMALIZE_GRAD = 2
GET_ALPHA, GET_WS, GET_VH, GET_PENALTY, GET_STATE, GET_CONTEXT = range(6)


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _f(t):
    assert t is None or t.dtype == torch.float32, "expected float32"
    return _p(t)


def _i(t):
    assert t is None or t.dtype == torch.int32, "expected int32"
    return _p(t)


class Context:
    """s2s_ctx bound to a device and (by default) to torch's current stream on it."""

    def __init__(self, device=0, stream="torch"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise S2SError("no CUDA device: libs2s_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream(self.device).cuda_stream if stream == "torch" else stream
            check(self.lib.s2s_ctx_create(int(device), C.c_void_p(int(s or 0)), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.s2s_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self.lib.s2s_ctx_synchronize(self.h))

    @property
    def launches(self):
        return int(self.lib.s2s_ctx_launch_count(self.h))

    KERNEL_CLASSES = ("gemm_tc", "gemm_simt", "gru_cluster", "dec_cluster_fwd", "dec_cluster_bwd", "attn_step", "lstm_cluster")

    def kernel_counts(self):
        """{kernel class: launches since creation} -- which path ran (tcgen05 vs SIMT GEMM, cluster loops vs per-step chains)"""
        return {k: int(self.lib.s2s_ctx_kernel_count(self.h, i)) for i, k in enumerate(self.KERNEL_CLASSES)}

    def set_graphs(self, enable):
        check(self.lib.s2s_ctx_set_graphs(self.h, int(bool(enable))))

    # caller-defined graphs: capture a sequence of library calls once, replay it with one launch.  The tensors the
    # captured calls wrote to must stay alive (and be the ones read afterwards); run the sequence eagerly once first.
    def graph_begin(self):
        check(self.lib.s2s_graph_begin(self.h))

    def graph_end(self):
        gid = C.c_int(-1)
        check(self.lib.s2s_graph_end(self.h, C.byref(gid)))
        return int(gid.value)

    def graph_launch(self, gid):
        check(self.lib.s2s_graph_launch(self.h, gid))

    def graph_destroy(self, gid):
        check(self.lib.s2s_graph_destroy(self.h, gid))

    PROF_CLASSES = ("attn_fwd", "attn_bwd", "attn_dvh", "gru_fwd", "gru_bwd", "gemm", "dense_small", "dec_fwd", "dec_bwd", "gemm_side")

    def profile(self, enable=True):
        check(self.lib.s2s_ctx_profile(self.h, int(bool(enable))))

    def profile_read(self):
        """{class: (ms, launches, algorithmic work)} since profile(True)"""
        n = len(self.PROF_CLASSES)
        ms = (C.c_double * n)(); cnt = (C.c_int64 * n)(); work = (C.c_double * n)()
        check(self.lib.s2s_ctx_profile_read(self.h, ms, cnt, work))
        return {k: (ms[i], cnt[i], work[i]) for i, k in enumerate(self.PROF_CLASSES)}

    def new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)


# ---- layout ----------------------------------------------------------------------------------------
def param_count(cfg):
    return int(_lib.load().s2s_param_count(C.byref(ModelCfg.from_dict(cfg))))


def param_segments(cfg):
    buf = (C.c_int64 * (3 * 128))()
    n = _lib.load().s2s_param_segments(C.byref(ModelCfg.from_dict(cfg)), buf, 128)
    return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]


def decoder_param_offset(cfg):
    return int(_lib.load().s2s_decoder_param_offset(C.byref(ModelCfg.from_dict(cfg))))


# ---- dense ------------------------------------------------------------------------------------------
def gemm(ctx, A, B, tA=False, tB=False, alpha=1.0, beta=0.0, C_out=None, bias=None, impl=0):
    M = A.shape[1] if tA else A.shape[0]
    K = A.shape[0] if tA else A.shape[1]
    N = B.shape[0] if tB else B.shape[1]
    if C_out is None:
        C_out = ctx.zeros(M, N)
    def ptr(t):   # row-strided 2-D views are allowed here (leading dimension = stride(0))
        assert t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1
        return C.c_void_p(t.data_ptr())
    check(ctx.lib.s2s_gemm_f32(ctx.h, impl, int(tA), int(tB), M, N, K, alpha, ptr(A), A.stride(0), ptr(B), B.stride(0), beta,
                               ptr(C_out), C_out.stride(0), _f(bias)))
    return C_out


def tconv_zb_forward(ctx, x, W):
    rows = x.numel() // x.shape[-1]
    y = ctx.new(*x.shape[:-1], W.shape[0])
    check(ctx.lib.s2s_tconv_zb_forward(ctx.h, _f(x), rows, x.shape[-1], _f(W), W.shape[0], _f(y)))
    return y


def tconv_zb_backward(ctx, x, W, dy, dW=None, scale=1.0, need_dx=True):
    rows = x.numel() // x.shape[-1]
    dx = ctx.new(*x.shape) if need_dx else None
    check(ctx.lib.s2s_tconv_zb_backward(ctx.h, _f(x), rows, x.shape[-1], _f(W), W.shape[0], _f(dy), _f(dx), _f(dW), scale))
    return dx


# ---- GRU sequence ---------------------------------------------------------------------------------------
def gru_seq_forward(ctx, W, x, lengths=None, ndir=1, reverse=False):
    """W: [ndir*3, H, H+Din] (z, r, h~ per direction); x [B, Lmax, Din] -> y [B, Lmax, ndir*H], save"""
    B, L, Din = x.shape
    H = W.shape[-2]
    assert W.shape[-1] == H + Din and W.numel() == ndir * 3 * H * (H + Din)
    y = ctx.new(B, L, ndir * H)
    save = ctx.new(int(ctx.lib.s2s_gru_seq_save_floats(B, L, H, ndir)))
    check(ctx.lib.s2s_gru_seq_forward(ctx.h, _f(W), Din, H, ndir, int(reverse), _f(x), Din, _i(lengths), B, L, _f(y), _f(save)))
    return y, save


def gru_seq_backward(ctx, W, x, y, save, dy, lengths=None, ndir=1, reverse=False, dW=None):
    B, L, Din = x.shape
    H = W.shape[-2]
    if dW is None:
        dW = torch.zeros_like(W)
    dx = ctx.new(B, L, Din)
    check(ctx.lib.s2s_gru_seq_backward(ctx.h, _f(W), _f(dW), Din, H, ndir, int(reverse), _f(x), Din, _i(lengths), B, L,
                                       _f(y), _f(save), _f(dy), _f(dx)))
    return dx, dW


def gru_step_forward(ctx, W, x, hprev=None):
    B, Din = x.shape
    H = W.shape[-2]
    hn = ctx.new(B, H); gates = ctx.new(B, 3 * H)
    check(ctx.lib.s2s_gru_step_forward(ctx.h, _f(W), Din, H, _f(x), _f(hprev), B, _f(hn), _f(gates)))
    return hn, gates


def gru_step_backward(ctx, W, x, hprev, gates, dhn, dW=None):
    B, Din = x.shape
    H = W.shape[-2]
    if dW is None:
        dW = torch.zeros_like(W)
    dx = ctx.new(B, Din); dhp = ctx.new(B, H)
    check(ctx.lib.s2s_gru_step_backward(ctx.h, _f(W), _f(dW), Din, H, _f(x), _f(hprev), B, _f(gates), _f(dhn), _f(dx), _f(dhp)))
    return dx, dhp, dW


def dropout_mask(ctx, shape, p, seed=0, out=None):
    m = ctx.new(*shape) if out is None else out
    check(ctx.lib.s2s_dropout_mask(ctx.h, p, seed, m.numel(), _f(m)))
    return m


def lstm_step_forward(ctx, P, x, H, hprev=None, cprev=None, peepholes=False):
    B, Din = x.shape
    hn = ctx.new(B, H); cn = ctx.new(B, H); acts = ctx.new(B, 4 * H)
    check(ctx.lib.s2s_lstm_step_forward(ctx.h, _f(P), Din, H, int(peepholes), _f(x), _f(hprev), _f(cprev), B, _f(hn), _f(cn), _f(acts)))
    return hn, cn, acts


def lstm_step_backward(ctx, P, x, H, hprev, cprev, acts, cnext, dhn, dcn=None, peepholes=False, dP=None):
    B, Din = x.shape
    if dP is None:
        dP = torch.zeros_like(P)
    dx = ctx.new(B, Din); dhp = ctx.new(B, H); dcp = ctx.new(B, H)
    check(ctx.lib.s2s_lstm_step_backward(ctx.h, _f(P), _f(dP), Din, H, int(peepholes), _f(x), _f(hprev), _f(cprev), B, _f(acts), _f(cnext),
                                         _f(dhn), _f(dcn), _f(dx), _f(dhp), _f(dcp)))
    return dx, dhp, dcp, dP


# ---- LSTM sequence --------------------------------------------------------------------------------------
def lstm_param_count(din, H, peepholes):
    return int(_lib.load().s2s_lstm_param_count(din, H, int(peepholes)))


def lstm_segments(din, H, peepholes):
    """[(name, (weight offset, rows, cols), bias offset)] of the flat LSTM parameter block, in the order the reference module's
    parameters() yields (LSTM.lua:25-36): per gate i, f, g, o: Linear(in,out), Linear(out,out) [, peephole Linear(out,out)], each with bias"""
    segs, o = [], 0
    for gi, g in enumerate("ifgo"):
        segs.append((g + ".x", (o, H, din), o + H * din)); o += H * din + H
        segs.append((g + ".h", (o, H, H), o + H * H)); o += H * H + H
        if peepholes and g != "g":
            segs.append((g + ".c", (o, H, H), o + H * H)); o += H * H + H
    assert o == lstm_param_count(din, H, peepholes)
    return segs


def lstm_seq_forward(ctx, P, x, H, peepholes=False, lengths=None, reverse=False):
    B, L, Din = x.shape
    assert P.numel() == lstm_param_count(Din, H, peepholes)
    y = ctx.new(B, L, H)
    save = ctx.new(int(ctx.lib.s2s_lstm_seq_save_floats(B, L, H)))
    check(ctx.lib.s2s_lstm_seq_forward(ctx.h, _f(P), Din, H, int(peepholes), int(reverse), _f(x), Din, _i(lengths), B, L, _f(y), _f(save)))
    return y, save


def lstm_seq_backward(ctx, P, x, y, save, dy, H, peepholes=False, lengths=None, reverse=False, dP=None):
    B, L, Din = x.shape
    if dP is None:
        dP = torch.zeros_like(P)
    dx = ctx.new(B, L, Din)
    check(ctx.lib.s2s_lstm_seq_backward(ctx.h, _f(P), _f(dP), Din, H, int(peepholes), int(reverse), _f(x), Din, _i(lengths), B, L,
                                        _f(y), _f(save), _f(dy), _f(dx)))
    return dx, dP


# ---- attention decoder ---------------------------------------------------------------------------------
def attention_forward(ctx, cfg, P, h, labels, lengths=None, tlens=None, dropmask=None, lam=0.0):
    B, L, A = h.shape
    T = labels.shape[1]
    logp = ctx.new(B, T, cfg["V"])
    check(ctx.lib.s2s_attention_forward(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(h), _i(lengths), B, L, _i(labels), _i(tlens), T,
                                        _f(dropmask), lam, _f(logp)))
    return logp


def attention_backward(ctx, cfg, P, G, h, labels, dlogp, lengths=None, tlens=None, dropmask=None, lam=0.0):
    B, L, A = h.shape
    T = labels.shape[1]
    dh = ctx.new(B, L, A)
    check(ctx.lib.s2s_attention_backward(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(G), _f(h), _i(lengths), B, L, _i(labels),
                                         _i(tlens), T, _f(dropmask), lam, _f(dlogp), _f(dh)))
    return dh


def attention_get(ctx, what, shape):
    out = ctx.new(*shape)
    check(ctx.lib.s2s_attention_get(ctx.h, what, _f(out)))
    return out


def attention_step(ctx, cfg, P, h, Vh, yprev=None, alpha_prev=None, s_prev=None, lengths=None):
    B, L, A = h.shape
    alpha = ctx.new(B, L); s = ctx.new(B, cfg["ST"]); logp = ctx.new(B, cfg["V"])
    check(ctx.lib.s2s_attention_step(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(h), _f(Vh), _i(lengths), B, L, _i(yprev),
                                     _f(alpha_prev), _f(s_prev), _f(alpha), _f(s), _f(logp)))
    return alpha, s, logp


def beam_search(ctx, cfg, P, h, eos, beam=5, maxlen=None):
    """Attention:BeamSearch for one utterance h [L, A]; returns (labels list, total log-prob)"""
    L = h.shape[0]
    maxlen = maxlen or L
    out = (C.c_int * (maxlen + 2))()
    n = C.c_int(0)
    lp = C.c_float(0)
    check(ctx.lib.s2s_beam_search(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(h), L, int(eos), int(beam), int(maxlen),
                                  out, C.byref(n), C.byref(lp)))
    return [out[i] for i in range(n.value)], float(lp.value)


# ---- whole model ----------------------------------------------------------------------------------------
def model_forward(ctx, cfg, P, X, labels, lengths=None, tlens=None, dropmask=None, lam=0.0, flags=0, want_logp=True):
    B, L, D = X.shape
    T = labels.shape[1]
    nll = ctx.new(B)
    logp = ctx.new(B, T, cfg["V"]) if want_logp else None
    check(ctx.lib.s2s_model_forward(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(X), _i(lengths), B, L, _i(labels), _i(tlens), T,
                                    _f(dropmask), lam, flags, _f(nll), _f(logp)))
    return nll, logp


def model_fwdbwd(ctx, cfg, P, G, X, labels, lengths=None, tlens=None, dropmask=None, lam=0.0, flags=0, nll=None, logp=None, dX=None):
    """G is accumulated (zero it first, as autoencoder:zeroGradParameters() does)."""
    B, L, D = X.shape
    T = labels.shape[1]
    if nll is None:
        nll = ctx.new(B)
    check(ctx.lib.s2s_model_fwdbwd(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), _f(G), _f(X), _i(lengths), B, L, _i(labels), _i(tlens), T,
                                   _f(dropmask), lam, flags, _f(nll), _f(logp), _f(dX)))
    return nll


def model_annotations(ctx, B, L, A):
    out = ctx.new(B, L, A)
    check(ctx.lib.s2s_model_get_annotations(ctx.h, _f(out)))
    return out


# ---- noise / optimiser --------------------------------------------------------------------------------
def weightnoise_sample(ctx, w, sigma, eps=None, seed=0):
    out = torch.empty_like(w)
    check(ctx.lib.s2s_weightnoise_sample(ctx.h, _f(w), _f(eps), seed, sigma, w.numel(), _f(out)))
    return out


def awn_sample(ctx, weight, eps=None, seed=0, out=None):
    n = weight.numel() // 2
    if out is None:
        out = ctx.new(n)
    check(ctx.lib.s2s_awn_sample(ctx.h, _f(weight), _f(eps), seed, n, _f(out)))
    return out


def awn_forward(ctx, weight, lam, nll):
    L = C.c_double(0)
    check(ctx.lib.s2s_awn_forward(ctx.h, _f(weight), weight.numel() // 2, lam, nll, C.byref(L)))
    return float(L.value)


def awn_accgrad(ctx, weight, g, lam, out=None):
    gw = torch.empty_like(weight) if out is None else out
    check(ctx.lib.s2s_awn_accgrad(ctx.h, _f(weight), _f(g), g.numel(), lam, _f(gw)))
    return gw


def grad_finalize(ctx, g, p, batch, maxnorm, wd=0.0, noise=None, seed=0, noise_sigma=0.0, want_norm=True):
    nrm = C.c_double(0)
    check(ctx.lib.s2s_grad_finalize(ctx.h, _f(g), _f(p), g.numel(), batch, maxnorm, wd, _f(noise), seed, noise_sigma,
                                    C.byref(nrm) if want_norm else None))
    return float(nrm.value) if want_norm else None


def adadelta(ctx, x, g, v, a, rho=0.95, eps=1e-8):
    check(ctx.lib.s2s_adadelta(ctx.h, _f(x), _f(g), _f(v), _f(a), x.numel(), rho, eps))


def orthogonalize(ctx, W, bias=None):
    """TrainUtils.orthogonalize (TrainUtils.lua:5-26) in place on a 2-D weight (and its bias, appended as a column)."""
    assert W.dim() == 2
    check(ctx.lib.s2s_orthogonalize(ctx.h, _f(W), W.shape[0], W.shape[1], _f(bias)))
    return W


def rownorm_constraint(ctx, W, maxval=1.0):
    flag = C.c_int(0)
    check(ctx.lib.s2s_rownorm_constraint(ctx.h, _f(W), W.shape[0], W.shape[1], maxval, C.byref(flag)))
    return flag.value


def model_rownorm_constraint(ctx, cfg, P, maxval=1.0):
    flag = C.c_int(0)
    check(ctx.lib.s2s_model_rownorm_constraint(ctx.h, C.byref(ModelCfg.from_dict(cfg)), _f(P), maxval, C.byref(flag)))
    return flag.value


# ---- kernel-level hooks (microbenchmarks) -------------------------------------------------------------
def attn_step_forward(ctx, Vh, h, q, w, lengths=None, alpha=None, c=None):
    B, L, S = Vh.shape
    A = h.shape[2]
    alpha = ctx.new(B, L) if alpha is None else alpha
    c = ctx.new(B, A) if c is None else c
    check(ctx.lib.s2s_attn_step_forward(ctx.h, _f(Vh), _f(h), _f(q), _f(w), _i(lengths), B, L, S, A, _f(alpha), _f(c)))
    return alpha, c


def attn_step_backward(ctx, Vh, h, q, w, alpha, dc, dalpha_in=None, lengths=None, dq=None, de=None):
    B, L, S = Vh.shape
    A = h.shape[2]
    dq = ctx.new(B, S) if dq is None else dq
    de = ctx.new(B, L) if de is None else de
    check(ctx.lib.s2s_attn_step_backward(ctx.h, _f(Vh), _f(h), _f(q), _f(w), _i(lengths), B, L, S, A, _f(alpha), _f(dc), _f(dalpha_in),
                                         _f(dq), _f(de)))
    return dq, de


# ---- VGG front-end (librispeech/model_vgg.lua:23-54) ----------------------------------------------------------
VGG_LIBRISPEECH = dict(C1=64, C2=128, HID=2048, OUT=512)


def vgg_param_count(cfg, F):
    c = _lib.VggCfg.of(cfg)
    return int(_lib.load().s2s_vgg_param_count(C.byref(c), F))


def vgg_forward(ctx, cfg, P, X):
    """X [B, 3, T, F] -> annotations [B, (T-8)//2, OUT]"""
    c = _lib.VggCfg.of(cfg)
    B, _, T, F = X.shape
    h = ctx.new(B, (T - 8) // 2, c.OUT)
    check(ctx.lib.s2s_vgg_forward(ctx.h, C.byref(c), _f(P), _f(X), B, T, F, _f(h)))
    return h


def vgg_backward(ctx, cfg, P, X, dh, dP=None, need_dx=False):
    c = _lib.VggCfg.of(cfg)
    B, _, T, F = X.shape
    if dP is None:
        dP = torch.zeros_like(P)
    dX = torch.empty_like(X) if need_dx else None
    check(ctx.lib.s2s_vgg_backward(ctx.h, C.byref(c), _f(P), _f(dP), B, T, F, _f(dh), _f(dX)))
    return dP, dX


def labels_from_onehot(ctx, onehot):
    """labelmask [..., V] (float one-hot, timit/timit.lua:262) -> int32 labels [...]; -1 for all-zero rows"""
    V = onehot.shape[-1]
    labels = torch.empty(onehot.shape[:-1], dtype=torch.int32, device=onehot.device)
    check(ctx.lib.s2s_labels_from_onehot(ctx.h, _f(onehot), labels.numel(), V, _i(labels)))
    return labels


def onehot(ctx, labels, V):
    out = torch.empty(*labels.shape, V, dtype=torch.float32, device=labels.device)
    check(ctx.lib.s2s_onehot(ctx.h, _i(labels), labels.numel(), V, _f(out)))
    return out


def nll_grad_seed(ctx, logp, labels, tlens=None, flags=0, nll=None, dlogp=None, want_grad=True):
    """per-utterance NLL [B] and the gradient seed dlogp = -labelmask [B,T,V] (timit/timit.lua:262-282)"""
    B, T, V = logp.shape
    if nll is None:
        nll = ctx.new(B)
    if dlogp is None and want_grad:
        dlogp = ctx.new(B, T, V)
    check(ctx.lib.s2s_nll_grad_seed(ctx.h, _f(logp), _i(labels), _i(tlens), B, T, V, flags, _f(nll), _f(dlogp)))
    return nll, dlogp


def conv3_forward(ctx, x, Ww, Wp, bias=None, relu=False):
    """implicit 3x3 convolution test hook: x [Mg, C] (pixels of a grid with row pitch Ww), Wp [N, 9*C] -> [Mg, N]"""
    Mg, Cc = x.shape
    N = Wp.shape[0]
    out = ctx.new(Mg, N)
    check(ctx.lib.s2s_conv3_forward(ctx.h, _f(x), Mg, Ww, Cc, _f(Wp), _f(bias), N, _f(out), int(relu)))
    return out


def conv3_dgrad(ctx, dout, Ww, WpT):
    Mg, N = dout.shape
    Cc = WpT.shape[0]
    din = ctx.new(Mg, Cc)
    check(ctx.lib.s2s_conv3_dgrad(ctx.h, _f(dout), Mg, Ww, N, _f(WpT), Cc, _f(din)))
    return din


def conv3_wgrad(ctx, dout, x, Ww, dWp):
    Mg, N = dout.shape
    check(ctx.lib.s2s_conv3_wgrad(ctx.h, _f(dout), _f(x), Mg, Ww, N, x.shape[1], _f(dWp)))
    return dWp


def edit_distance(a, b):
    """WagnerFischer(a, b) of utils.lua:3-27 on two host label sequences."""
    import numpy as np
    from ._lib import load
    a = np.ascontiguousarray(a, dtype=np.int32); b = np.ascontiguousarray(b, dtype=np.int32)
    out = C.c_int(0)
    check(load().s2s_edit_distance(a.ctypes.data_as(C.c_void_p), a.size, b.ctypes.data_as(C.c_void_p), b.size, C.byref(out)))
    return int(out.value)


def attn_step_forward_loc(ctx, Vh, h, q, w, uw, alpha_prev, lengths=None, alpha=None, c=None):
    B, L, S = Vh.shape
    A = h.shape[2]
    alpha = ctx.new(B, L) if alpha is None else alpha
    c = ctx.new(B, A) if c is None else c
    check(ctx.lib.s2s_attn_step_forward_loc(ctx.h, _f(Vh), _f(h), _f(q), _f(w), _i(lengths), B, L, S, A, uw.shape[0], _f(uw), _f(alpha_prev),
                                            _f(alpha), _f(c)))
    return alpha, c


def attn_step_backward_loc(ctx, Vh, h, q, w, uw, alpha_prev, alpha, dc, dalpha_in=None, lengths=None, dq=None, de=None, dalpha_prev=None):
    B, L, S = Vh.shape
    A = h.shape[2]
    dq = ctx.new(B, S) if dq is None else dq
    de = ctx.new(B, L) if de is None else de
    dalpha_prev = ctx.new(B, L) if dalpha_prev is None else dalpha_prev
    check(ctx.lib.s2s_attn_step_backward_loc(ctx.h, _f(Vh), _f(h), _f(q), _f(w), _i(lengths), B, L, S, A, uw.shape[0], _f(uw), _f(alpha_prev),
                                             _f(alpha), _f(dc), _f(dalpha_in), _f(dq), _f(de), _f(dalpha_prev)))
    return dq, de, dalpha_prev


# ---- parameter initialisation (module:reset(stdv) rules) -------------------------------------------------
def segment_names(cfg):
    """Names of the flat-layout segments in order (see csrc/core.cu make_layout)."""
    names = [f"enc{l}{d}.W{g}" for l in range(cfg["NL"]) for d in ("f", "r") for g in ("z", "r", "h")]
    names += ["WV", "bV", "Ws", "bs"]
    if cfg["K"] > 0:
        names += ["WF", "bF", "U", "bU"]
    names += ["we", "be", "Wy", "by", "Wc", "bc", "Wj", "bj", "Gz", "Gr", "Gh", "Wm", "bm"]
    if cfg.get("MLP", 1) == 2:                     # librispeech/model_vgg.lua:78-79
        names += ["Wl", "bl", "Wm2", "bm2"]
    names += ["Wo", "bo"]
    return names


def init_params(cfg, seed=1234):
    """U(+-1/sqrt(fan_in)) for every weight and live bias, as reset() does in LinearZeroBias.lua:12-29 and
    TemporalConvolutionZeroBias.lua:21-35 (stock nn.Linear / nn.TemporalConvolution use the same bound for
    the bias); the dead biases of the ZeroBias convolutions stay 0.  Returns a float32 numpy vector."""
    import numpy as np
    rng = np.random.default_rng(seed)
    P = np.zeros(param_count(cfg), dtype=np.float32)
    fan_in = 1
    for (off, rows, cols), name in zip(param_segments(cfg), segment_names(cfg)):
        if cols > 1 or name == "we":
            fan_in = cfg["KF"] if name == "WF" else cols
            P[off:off + rows * cols] = rng.uniform(-1, 1, rows * cols) / np.sqrt(fan_in)
        elif name not in ("bV", "bU", "be"):
            P[off:off + rows] = rng.uniform(-1, 1, rows) / np.sqrt(fan_in)
    return P
