"""Parity ON the timed path: the exact configurations bench.py measures (BASELINE.json configs[1]: B = 32, L = 300,
T = 50, 3 x biGRU-256, S = A = 512, ST = 256; content attention = `cfg2`, location-aware K = 16 / k = 10 = `cfg2loc`),
through s2s_model_fwdbwd with CUDA graphs ON and the default overlap of the weight-gradient GEMMs, against the float64
CPU oracle run per utterance like timit/timit.lua:240-289.  Three calls on one context with the same pointers: call 1
is eager, call 2 is captured and replayed once, call 3 is the REPLAYED graph -- all three must match the oracle.
The kernel-class counters prove which kernels ran: tcgen05 GEMMs, the persistent GRU clusters, the decoder cluster loops.
Bound: <= 1e-4 relative (fp32), the bound north_star states."""
import os

import numpy as np
import pytest
import torch

from oracle.oracle import init_params, segment_names
from tests.util import dev, make_batch, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4

CFG2 = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=0, KF=10, M=64, MW=7)
B, L, T = 32, 300, 50


def _seg_errs(cfg, orc, G, Gref):
    out = {}
    for (off, rows, cols), name in zip(orc.param_segments(cfg), segment_names(cfg)):
        a, b = G[off:off + rows * cols], Gref[off:off + rows * cols]
        scale = max(np.abs(b).max(), 1e-6 * np.abs(Gref).max())
        out[name] = float(np.abs(a - b).max() / scale)
    return out


def _threads():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 32))
    except Exception:
        return 8


def _penalty_margin(ref, lengths, tlens):
    """smallest |sum_l (L-l)(alpha_t - alpha_{t-1})| / (L/2) over all utterances and steps: the argument of the
    MonotonicAlignment max(0, .) (MonotonicAlignment.lua:27-41), whose SIGN decides whether the +-lambda (L+1-l) gradient
    is injected.  With freshly initialised weights alpha is nearly uniform and the argument is below fp32 resolution
    (1e-8 relative: the float32 and float64 ORACLES then disagree by 80% on dW_s), so the penalty case runs on weights
    scaled to give peaked, moving alignments and asserts its own decision margin."""
    m = np.inf
    for b in range(len(lengths)):
        Lb, Tb = int(lengths[b]), int(tlens[b])
        if Tb < 2:
            continue
        w = (ref["alpha"][b, :Tb, :Lb] * (Lb - np.arange(Lb))).sum(1)
        m = min(m, float(np.abs(w[1:] - w[:-1]).min() / (Lb / 2)))
    return m


@pytest.mark.parametrize("name,extra,ragged,lam,scale", [("cfg2", {}, False, 0.0, 1.0), ("cfg2loc", dict(K=16), False, 0.0, 1.0),
                                                         ("cfg2loc-ragged", dict(K=16), True, 0.0, 1.0),
                                                         ("cfg2loc-ragged-penalty", dict(K=16), True, 0.02, 4.0)])
def test_timed_configuration_matches_oracle(s2s, orc64, name, extra, ragged, lam, scale):
    cfg = dict(CFG2, **extra)
    P = init_params(cfg, seed=1234, dtype=np.float64, oracle=orc64) * scale
    X, lengths, labels, tlens = make_batch(cfg, B, L, T, seed=1000, ragged=ragged)
    ref = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, lam=lam, normalize_nll=True, nthreads=_threads())
    if lam != 0.0:
        assert _penalty_margin(ref, lengths, tlens) > 5e-6
    ctx = s2s.Context(0)
    try:
        Pd = dev(P, torch.float32); G = torch.zeros_like(Pd)
        Xd, ld, yd, td = dev(X), dev(lengths), dev(labels), dev(tlens)
        logp = ctx.new(B, T, cfg["V"]); dX = ctx.new(B, L, cfg["D"]); nll = ctx.new(B)
        k0 = ctx.kernel_counts()
        for call in range(3):                     # eager, captured + first replay, replayed graph
            G.zero_(); logp.zero_(); dX.zero_(); nll.zero_()
            s2s.model_fwdbwd(ctx, cfg, Pd, G, Xd, yd, lengths=ld, tlens=td, lam=lam, flags=s2s.NORMALIZE_NLL, nll=nll, logp=logp, dX=dX)
            torch.cuda.synchronize()
            assert rel_err(nll.cpu().numpy(), ref["nll"]) < TOL, (name, call)
            lp = logp.cpu().numpy(); dx = dX.cpu().numpy()
            for b in range(B):
                Lb, Tb = lengths[b], tlens[b]
                assert rel_err(lp[b, :Tb], ref["logp"][b, :Tb]) < TOL, (name, call, b)
                assert rel_err(dx[b, :Lb], ref["dX"][b, :Lb]) < TOL, (name, call, b)
            errs = _seg_errs(cfg, orc64, G.cpu().numpy(), ref["G"])
            bad = {k: v for k, v in errs.items() if v > TOL}
            assert not bad, (name, call, bad)
        k1 = ctx.kernel_counts()
        d = {k: k1[k] - k0[k] for k in k1}
        # the combination that is timed: tcgen05 GEMMs in situ, persistent GRU clusters (3 layers x fwd/bwd), one decoder
        # cluster kernel per direction of time -- and NOT the per-step attention chain
        assert d["gemm_tc"] >= 3 * 20, d
        assert d["gru_cluster"] == 3 * 6, d
        assert d["dec_cluster_fwd"] == 3, d
        if os.environ.get("S2S_TEST_LOC_BWD_CLUSTER", "1") == "1" or (cfg["K"] == 0 and lam == 0.0):
            assert d["dec_cluster_bwd"] == 3 and d["attn_step"] == 0, d
    finally:
        ctx.close()


@pytest.mark.parametrize("extra,lam", [({}, 0.0), (dict(K=16), 0.0), (dict(K=16), 0.03)])
def test_cluster_decoder_backward_equals_per_step_chain(s2s, orc64, extra, lam):
    """dec_cluster_bwd_kernel vs the per-step chain (attn_bwd + dense launches) on identical seeded inputs at the
    benchmark shape: elementwise <= 1e-5 of the tensor's scale (both are fp32 evaluations of the same sums in a
    different order)."""
    cfg = dict(CFG2, NL=0, **extra)
    P = dev(init_params(cfg, seed=1234, dtype=np.float64, oracle=orc64), torch.float32)
    rng = np.random.default_rng(7)
    h = dev(rng.standard_normal((B, L, 512)) * 0.5, torch.float32)
    y = dev(rng.integers(0, cfg["V"] - 1, (B, T)).astype(np.int32))
    dlogp = dev(rng.standard_normal((B, T, cfg["V"])), torch.float32)
    res = {}
    old = os.environ.get("S2S_DEC_CLUSTER_BWD")
    # ONE forward, two backward passes over its saved state: the monotonicity penalty is gated by the sign of sum(cumsum alpha_t -
    # cumsum alpha_{t-1}) (MonotonicAlignment.lua:19-77), which at random initialisation is a sum of near-cancelling terms -- two forward
    # runs whose fp32 sums differ in the last bit can flip a gate, and that is a property of the function, not of the backward kernels
    ctx = s2s.Context(0)
    try:
        s2s.attention_forward(ctx, cfg, P, h, y, lam=lam)
        for mode in ("1", "0"):
            os.environ["S2S_DEC_CLUSTER_BWD"] = mode
            G = torch.zeros_like(P)
            k0 = ctx.kernel_counts()
            dh = s2s.attention_backward(ctx, cfg, P, G, h, y, dlogp, lam=lam)
            torch.cuda.synchronize()
            k1 = ctx.kernel_counts()
            if os.environ.get("S2S_TEST_LOC_BWD_CLUSTER", "1") == "1" or (cfg["K"] == 0 and lam == 0.0):
                assert (k1["dec_cluster_bwd"] - k0["dec_cluster_bwd"]) == (1 if mode == "1" else 0)
            res[mode] = (dh.cpu().numpy().copy(), G.cpu().numpy().copy())
    finally:
        ctx.close()
        if old is None:
            os.environ.pop("S2S_DEC_CLUSTER_BWD", None)
        else:
            os.environ["S2S_DEC_CLUSTER_BWD"] = old
    # with the penalty, d alpha carries +-lambda (L - l) terms of size ~10 beside a 1e-3 signal and the softmax backward cancels them:
    # the two evaluation orders then differ by the cancellation noise, not by 1e-7
    tol = 1e-5 if lam == 0.0 else 1e-3
    assert rel_err(res["1"][0], res["0"][0]) < tol
    segs = list(zip(orc64.param_segments(cfg), segment_names(cfg)))
    for (off, rows, cols), name in segs:
        a, b = res["1"][1][off:off + rows * cols], res["0"][1][off:off + rows * cols]
        scale = max(np.abs(b).max(), 1e-6 * np.abs(res["0"][1]).max())
        assert np.abs(a - b).max() / scale < 2 * tol, name


def test_normalize_grad_flag(s2s, gctx, orc64):
    """S2S_NORMALIZE_GRAD scales the gradient seed by 1/T_b independently of S2S_NORMALIZE_NLL (timit.lua:268-281)."""
    cfg = dict(D=13, H=128, NL=2, S=128, ST=64, V=11, K=0, KF=4, M=8, MW=3)
    Bs, Ls, Ts = 4, 30, 6
    P = init_params(cfg, seed=3, dtype=np.float64, oracle=orc64) * 1.5
    X, lengths, labels, tlens = make_batch(cfg, Bs, Ls, Ts, seed=5)
    for nn_, ng in ((False, True), (True, True), (True, False)):
        ref = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, normalize_nll=nn_, normalize_grad=ng, nthreads=2)
        Pd = dev(P, torch.float32); G = torch.zeros_like(Pd)
        flags = (s2s.NORMALIZE_NLL if nn_ else 0) | (s2s.NORMALIZE_GRAD if ng else 0)
        nll = s2s.model_fwdbwd(gctx, cfg, Pd, G, dev(X), dev(labels), lengths=dev(lengths), tlens=dev(tlens), flags=flags)
        assert rel_err(nll.cpu().numpy(), ref["nll"]) < TOL
        assert rel_err(G.cpu().numpy(), ref["G"]) < TOL
