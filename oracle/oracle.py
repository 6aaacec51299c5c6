"""ctypes binding of the CPU oracle (oracle/s2s_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: unpinned beyond the four notebook known-answer cells (see s2s_oracle.c header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))

CFG_KEYS = ("D", "H", "NL", "S", "ST", "V", "K", "KF", "M", "MW", "MLP")
CFG_DEFAULTS = {"MLP": 1}     # MLP: 1 = Maxout-Linear (TIMIT models), 2 = Maxout-Linear-Maxout-Linear (librispeech/model_vgg.lua:76-80)
# timit/model_chorowski_baseline.lua:14-46 defaults
CHOROWSKI_TIMIT = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=0, KF=10, M=64, MW=7)


def build(force=False):
    """Compile both precisions of the oracle with the committed Makefile."""
    outs = [os.path.join(_DIR, f"liboracle_f{b}.so") for b in (32, 64)]
    src = os.path.join(_DIR, "s2s_oracle.c")
    if force or any(not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(src) for o in outs):
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return outs


def cfg_array(cfg):
    return (C.c_int * len(CFG_KEYS))(*[int(cfg.get(k, CFG_DEFAULTS.get(k))) for k in CFG_KEYS])


class Oracle:
    """One precision of the oracle.  All arrays are C-contiguous numpy arrays of self.dtype."""

    def __init__(self, precision="f32"):
        path = os.path.join(_DIR, f"liboracle_{precision}.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.creal = C.c_float if precision == "f32" else C.c_double
        assert self.lib.orc_sizeof_real() == np.dtype(self.dtype).itemsize
        self.lib.orc_param_count.restype = C.c_int64
        self.lib.orc_lstm_param_count.restype = C.c_int64
        self.lib.orc_awn_forward.restype = C.c_double
        self.lib.orc_grad_finalize.restype = C.c_double
        self.lib.orc_addbias_gradbias.restype = self.creal

    # -- helpers -------------------------------------------------------------------------
    def a(self, x):
        return np.ascontiguousarray(x, dtype=self.dtype)

    def p(self, x):
        if x is None:
            return None
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data_as(C.c_void_p)

    @staticmethod
    def ip(x):
        if x is None:
            return None
        assert x.dtype == np.int32 and x.flags["C_CONTIGUOUS"]
        return x.ctypes.data_as(C.c_void_p)

    def r(self, v):
        return self.creal(v)

    # -- layout --------------------------------------------------------------------------
    def param_count(self, cfg):
        return int(self.lib.orc_param_count(cfg_array(cfg)))

    def param_segments(self, cfg):
        buf = np.zeros(3 * 64, dtype=np.int64)
        n = self.lib.orc_param_segments(cfg_array(cfg), buf.ctypes.data_as(C.c_void_p))
        return buf[: 3 * n].reshape(n, 3).copy()

    # -- primitives pinned by notebook cells ----------------------------------------------
    def tconv_forward(self, x, W, b, kW, dW=1):
        x = self.a(x); W = self.a(W); L, inp = x.shape; out = W.shape[0]
        b = None if b is None else self.a(b)
        y = np.zeros(((L - kW) // dW + 1, out), dtype=self.dtype)
        self.lib.orc_tconv_forward(self.p(x), L, inp, self.p(W), self.p(b), out, kW, dW, self.p(y))
        return y

    def padding(self, x, pad):
        x = self.a(x); L, F = x.shape
        y = np.zeros((L + abs(pad), F), dtype=self.dtype)
        self.lib.orc_padding(self.p(x), L, F, pad, self.p(y))
        return y

    def mm(self, a, b):
        a = self.a(a); b = self.a(b)
        c = np.zeros((a.shape[0], b.shape[1]), dtype=self.dtype)
        self.lib.orc_mm(self.p(a), self.p(b), a.shape[0], a.shape[1], b.shape[1], self.p(c))
        return c

    def addbias_gradbias(self, go):
        go = self.a(go).reshape(-1)
        return float(self.lib.orc_addbias_gradbias(self.p(go), go.size))

    # -- GRU / LSTM ------------------------------------------------------------------------
    def gru_step_forward(self, Wz, Wr, Wh, x, hp):
        Wz, Wr, Wh, x, hp = map(self.a, (Wz, Wr, Wh, x, hp)); out = Wz.shape[0]; inp = Wz.shape[1] - out
        hn, z, r, hc = (np.zeros(out, dtype=self.dtype) for _ in range(4))
        self.lib.orc_gru_step_forward(self.p(Wz), self.p(Wr), self.p(Wh), inp, out, self.p(x), self.p(hp), self.p(hn), self.p(z), self.p(r), self.p(hc))
        return hn, z, r, hc

    def gru_step_backward(self, Wz, Wr, Wh, x, hp, z, r, hc, dhn):
        Wz, Wr, Wh, x, hp, z, r, hc, dhn = map(self.a, (Wz, Wr, Wh, x, hp, z, r, hc, dhn)); out = Wz.shape[0]; inp = Wz.shape[1] - out
        dWz, dWr, dWh = (np.zeros_like(Wz) for _ in range(3)); dx = np.zeros(inp, dtype=self.dtype); dhp = np.zeros(out, dtype=self.dtype)
        self.lib.orc_gru_step_backward(self.p(Wz), self.p(Wr), self.p(Wh), self.p(dWz), self.p(dWr), self.p(dWh), inp, out,
                                       self.p(x), self.p(hp), self.p(z), self.p(r), self.p(hc), self.p(dhn), self.p(dx), self.p(dhp))
        return dx, dhp, dWz, dWr, dWh

    def gru_seq_forward(self, Wz, Wr, Wh, x, reverse=False):
        Wz, Wr, Wh, x = map(self.a, (Wz, Wr, Wh, x)); out = Wz.shape[0]; inp = Wz.shape[1] - out; L = x.shape[0]
        y = np.zeros((L, out), dtype=self.dtype); gates = np.zeros((L, 3 * out), dtype=self.dtype)
        self.lib.orc_gru_seq_forward(self.p(Wz), self.p(Wr), self.p(Wh), inp, out, self.p(x), L, int(reverse), self.p(y), self.p(gates))
        return y, gates

    def gru_seq_backward(self, Wz, Wr, Wh, x, y, gates, dy, reverse=False):
        Wz, Wr, Wh, x, y, gates, dy = map(self.a, (Wz, Wr, Wh, x, y, gates, dy)); out = Wz.shape[0]; inp = Wz.shape[1] - out; L = x.shape[0]
        dWz, dWr, dWh = (np.zeros_like(Wz) for _ in range(3)); dx = np.zeros_like(x)
        self.lib.orc_gru_seq_backward(self.p(Wz), self.p(Wr), self.p(Wh), self.p(dWz), self.p(dWr), self.p(dWh), inp, out,
                                      self.p(x), L, int(reverse), self.p(y), self.p(gates), self.p(dy), self.p(dx))
        return dx, dWz, dWr, dWh

    def lstm_param_count(self, inp, out, peep):
        return int(self.lib.orc_lstm_param_count(inp, out, int(peep)))

    def lstm_seq_forward(self, P, inp, out, peep, x, reverse=False):
        P, x = self.a(P), self.a(x); L = x.shape[0]
        y = np.zeros((L, out), dtype=self.dtype); c = np.zeros((L, out), dtype=self.dtype); acts = np.zeros((L, 4 * out), dtype=self.dtype)
        self.lib.orc_lstm_seq_forward(self.p(P), inp, out, int(peep), self.p(x), L, int(reverse), self.p(y), self.p(c), self.p(acts))
        return y, c, acts

    def lstm_seq_backward(self, P, inp, out, peep, x, y, c, acts, dy, reverse=False):
        P, x, y, c, acts, dy = map(self.a, (P, x, y, c, acts, dy)); L = x.shape[0]
        dP = np.zeros_like(P); dx = np.zeros_like(x)
        self.lib.orc_lstm_seq_backward(self.p(P), self.p(dP), inp, out, int(peep), self.p(x), L, int(reverse), self.p(y), self.p(c), self.p(acts), self.p(dy), self.p(dx))
        return dx, dP

    # -- attention decoder ------------------------------------------------------------------
    def attention_forward(self, cfg, P, h, labels, lam=0.0, dropmask=None):
        P, h = self.a(P), self.a(h); labels = np.ascontiguousarray(labels, dtype=np.int32); L, T = h.shape[0], labels.shape[0]
        dm = None if dropmask is None else self.a(dropmask)
        A = 2 * cfg["H"]
        out = dict(logp=np.zeros((T, cfg["V"]), self.dtype), alpha=np.zeros((T, L), self.dtype), s=np.zeros((T, cfg["ST"]), self.dtype),
                   c=np.zeros((T, A), self.dtype), q=np.zeros((T, cfg["S"]), self.dtype), Vh=np.zeros((L, cfg["S"]), self.dtype), pen=np.zeros(T, self.dtype))
        self.lib.orc_attention_forward(cfg_array(cfg), self.p(P), self.r(lam), self.p(h), L, self.ip(labels), T, self.p(dm),
                                       *[self.p(out[k]) for k in ("logp", "alpha", "s", "c", "q", "Vh", "pen")])
        return out

    def attention_backward(self, cfg, P, h, labels, dlogp, lam=0.0, dropmask=None):
        P, h, dlogp = self.a(P), self.a(h), self.a(dlogp); labels = np.ascontiguousarray(labels, dtype=np.int32); L, T = h.shape[0], labels.shape[0]
        dm = None if dropmask is None else self.a(dropmask)
        G = np.zeros_like(P); dh = np.zeros_like(h)
        self.lib.orc_attention_backward(cfg_array(cfg), self.p(P), self.p(G), self.r(lam), self.p(h), L, self.ip(labels), T, self.p(dm), self.p(dlogp), self.p(dh))
        return G, dh

    # -- whole model --------------------------------------------------------------------------
    def model_fwdbwd(self, cfg, P, X, lengths, labels, tlens, lam=0.0, dropmask=None, normalize_nll=False, normalize_grad=False,
                     backward=True, nthreads=1, want=("logp", "alpha", "annot", "dX")):
        P, X = self.a(P), self.a(X); B, Lmax, D = X.shape
        labels = np.ascontiguousarray(labels, dtype=np.int32); Tmax = labels.shape[1]
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        tlens = None if tlens is None else np.ascontiguousarray(tlens, dtype=np.int32)
        dm = None if dropmask is None else self.a(dropmask)
        A = 2 * cfg["H"]
        G = np.zeros_like(P); nll = np.zeros(B, self.dtype)
        outs = dict(logp=np.zeros((B, Tmax, cfg["V"]), self.dtype) if "logp" in want else None,
                    alpha=np.zeros((B, Tmax, Lmax), self.dtype) if "alpha" in want else None,
                    annot=np.zeros((B, Lmax, A), self.dtype) if "annot" in want else None,
                    dX=np.zeros_like(X) if ("dX" in want and backward) else None)
        flags = (1 if normalize_nll else 0) | (2 if normalize_grad else 0) | (0 if backward else 4)
        self.lib.orc_model_fwdbwd(cfg_array(cfg), self.p(P), self.p(G), self.r(lam), self.p(X), self.ip(lengths), B, Lmax,
                                  self.ip(labels), self.ip(tlens), Tmax, self.p(dm), flags, int(nthreads),
                                  self.p(nll), self.p(outs["logp"]), self.p(outs["alpha"]), self.p(outs["annot"]), self.p(outs["dX"]))
        outs["nll"] = nll; outs["G"] = G
        return outs

    def beam_search(self, cfg, P, h, eos, K=5, maxlen=None):
        P, h = self.a(P), self.a(h); L = h.shape[0]; maxlen = maxlen or L
        out = np.zeros(maxlen + 2, dtype=np.int32); lp = self.creal(0)
        n = self.lib.orc_beam_search(cfg_array(cfg), self.p(P), self.p(h), L, int(eos), int(K), int(maxlen), self.ip(out), C.byref(lp))
        return out[:n].copy(), float(lp.value)

    # -- noise / optimiser ----------------------------------------------------------------------
    def weightnoise_sample(self, w, eps, sigma):
        w, eps = self.a(w), self.a(eps); s = np.zeros_like(w)
        self.lib.orc_weightnoise_sample(self.p(w), self.p(eps), self.r(sigma), C.c_int64(w.size), self.p(s)); return s

    def awn_sample(self, weight, eps):
        weight, eps = self.a(weight), self.a(eps); n = eps.size; s = np.zeros(n, self.dtype)
        self.lib.orc_awn_sample(self.p(weight), self.p(eps), C.c_int64(n), self.p(s)); return s

    def awn_forward(self, weight, lam, nll):
        weight = self.a(weight)
        return float(self.lib.orc_awn_forward(self.p(weight), C.c_int64(weight.size // 2), C.c_double(lam), C.c_double(nll)))

    def awn_accgrad(self, weight, g, lam):
        weight, g = self.a(weight), self.a(g); gw = np.zeros_like(weight)
        self.lib.orc_awn_accgrad(self.p(weight), self.p(g), C.c_int64(g.size), C.c_double(lam), self.p(gw)); return gw

    def grad_finalize(self, g, p, batch, maxnorm, wd, noise=None, noise_sigma=0.0):
        """in place on g; returns pre-clip norm"""
        assert g.dtype == self.dtype and g.flags["C_CONTIGUOUS"]
        p = self.a(p); noise = None if noise is None else self.a(noise)
        return float(self.lib.orc_grad_finalize(self.p(g), self.p(p), C.c_int64(g.size), int(batch), C.c_double(maxnorm), C.c_double(wd), self.p(noise), C.c_double(noise_sigma)))

    def adadelta(self, x, g, v, a, rho=0.95, eps=1e-8):
        for t in (x, g, v, a):
            assert t.dtype == self.dtype and t.flags["C_CONTIGUOUS"]
        self.lib.orc_adadelta(self.p(x), self.p(g), self.p(v), self.p(a), C.c_int64(x.size), C.c_double(rho), C.c_double(eps))

    def rownorm_constraint(self, W, maxval=1.0):
        assert W.dtype == self.dtype and W.flags["C_CONTIGUOUS"] and W.ndim == 2
        return int(self.lib.orc_rownorm_constraint(self.p(W), C.c_int64(W.shape[0]), C.c_int64(W.shape[1]), C.c_double(maxval)))

    def max_threads(self):
        return int(self.lib.orc_max_threads())


def init_params(cfg, seed=1234, dtype=np.float32, oracle=None):
    """U(+-1/sqrt(fan_in)) per reset() rules (LinearZeroBias.lua:12-29, TemporalConvolutionZeroBias.lua:21-35;
    stock nn.Linear/TemporalConvolution use the same bound for bias); dead ZeroBias biases = 0.
    Counter-based numpy RNG -- the reference seeds nothing, so any fixed seed is as faithful as another."""
    o = oracle or Oracle("f64")
    segs = o.param_segments(cfg)
    n = o.param_count(cfg)
    rng = np.random.default_rng(seed)
    P = np.zeros(n, dtype=np.float64)
    K = cfg["K"]
    # identify bias segments that are dead (Vh bias, U bias, e bias): they follow WV, U, we.
    names = segment_names(cfg)
    fan_in = None
    for (off, rows, cols), name in zip(segs, names):
        if cols > 1 or name in ("we",):
            fan_in = cols
            if name == "WF":
                fan_in = cfg["KF"]  # kW * inputFrameSize(1)
            P[off:off + rows * cols] = rng.uniform(-1, 1, rows * cols) / np.sqrt(fan_in)
        else:  # bias
            if name in ("bV", "bU", "be"):
                continue
            P[off:off + rows] = rng.uniform(-1, 1, rows) / np.sqrt(fan_in)
    return P.astype(dtype)


def segment_names(cfg):
    names = []
    for l in range(cfg["NL"]):
        for d in ("f", "r"):
            for g in ("z", "r", "h"):
                names.append(f"enc{l}{d}.W{g}")
    names += ["WV", "bV", "Ws", "bs"]
    if cfg["K"] > 0:
        names += ["WF", "bF", "U", "bU"]
    names += ["we", "be", "Wy", "by", "Wc", "bc", "Wj", "bj", "Gz", "Gr", "Gh", "Wm", "bm"]
    if cfg.get("MLP", 1) == 2:                     # librispeech/model_vgg.lua:78-79
        names += ["Wl", "bl", "Wm2", "bm2"]
    names += ["Wo", "bo"]
    return names
