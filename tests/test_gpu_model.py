"""GPU parity of nn.Attention and the whole Chorowski model (forward, per-utterance NLL, every
gradient) against the CPU oracle run per utterance exactly like timit/timit.lua:240-289.
All calls go through the C ABI.  Bound: <= 1e-4 relative (fp32), stated by north_star."""
import numpy as np
import pytest
import torch

from oracle.oracle import init_params
from tests.util import dev, make_batch, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4

MID = dict(D=13, H=128, NL=2, S=128, ST=64, V=11, K=0, KF=4, M=8, MW=3)


def _seg_errs(cfg, orc, G, Gref):
    from oracle.oracle import segment_names
    out = {}
    for (off, rows, cols), name in zip(orc.param_segments(cfg), segment_names(cfg)):
        a, b = G[off:off + rows * cols], Gref[off:off + rows * cols]
        scale = max(np.abs(b).max(), 1e-6 * np.abs(Gref).max())
        out[name] = float(np.abs(a - b).max() / scale)
    return out


@pytest.mark.parametrize("K,KF,lam,drop,extra", [(0, 4, 0.0, False, {}), (0, 4, 0.0, True, {}), (3, 4, 0.0, False, {}), (4, 5, 0.0, False, {}),
                                                 (2, 10, 0.02, False, {}), (0, 4, 0.05, False, {}),
                                                 # librispeech/model_vgg.lua decoder: two Maxout stages; decoder-only parameter vector
                                                 (0, 4, 0.0, True, dict(MLP=2)), (0, 10, 0.0, False, dict(MLP=2, NL=0))])
def test_attention_module_matches_oracle(s2s, gctx, orc64, K, KF, lam, drop, extra):
    _check_attention_module(s2s, gctx, orc64, dict(MID, K=K, KF=KF, **extra), 4, 45, 7, lam, drop)


# the model's own decoder sizes (ST = 256, S = A = 512): the time loop runs as ONE persistent cluster kernel
# (csrc/decoder_cluster.cu), and so does the backward loop when there is no alpha carry (lam = 0); B picks the
# utterances-per-cluster variant, short utterances leave CTAs without frames
@pytest.mark.parametrize("B,L,T,lam,extra", [(3, 70, 6, 0.03, {}), (3, 70, 6, 0.0, {}), (9, 40, 5, 0.0, {}), (16, 33, 4, 0.02, dict(MLP=2)),
                                             (16, 33, 4, 0.0, dict(MLP=2)), (37, 24, 3, 0.0, {}), (2, 400, 3, 0.0, {}),
                                             # location-aware term inside the forward cluster kernel (filter halo across CTAs)
                                             (4, 70, 6, 0.0, dict(K=16, KF=10)), (7, 45, 5, 0.02, dict(K=3, KF=5)), (2, 300, 4, 0.0, dict(K=2, KF=4)),
                                             # every utterances-per-cluster variant of the location-aware kernels (BG = 3, 4, 5: a shared-memory
                                             # member read as float4 was 16-byte aligned only for some BG -- round-2 regression)
                                             (17, 40, 4, 0.0, dict(K=16, KF=10)), (23, 33, 3, 0.02, dict(K=3, KF=5)), (30, 24, 3, 0.0, dict(K=16, KF=10))])
def test_attention_module_cluster_decoder(s2s, gctx, orc64, B, L, T, lam, extra):
    cfg = dict(dict(D=13, H=256, NL=0, S=512, ST=256, V=11, K=0, KF=4, M=8, MW=3), **extra)
    _check_attention_module(s2s, gctx, orc64, cfg, B, L, T, lam, False, short=True)


def _check_attention_module(s2s, gctx, orc64, cfg, B, L, T, lam, drop, short=False):
    P = init_params(cfg, seed=5, dtype=np.float64, oracle=orc64) * 1.5
    rng = np.random.default_rng(1)
    A = 2 * cfg["H"]
    h = rng.standard_normal((B, L, A)) * 0.7
    _, lengths, labels, tlens = make_batch(cfg, B, L, T, seed=2)
    if short:
        lengths[1] = 5; lengths[B - 1] = 17
    dm = ((rng.random((B, T, cfg["ST"] + A)) > 0.5) / 0.5) if drop else None
    dlogp = rng.standard_normal((B, T, cfg["V"]))
    for b in range(B):
        h[b, lengths[b]:] = 0; dlogp[b, tlens[b]:] = 0
    Pd, hd = dev(P, torch.float32), dev(h, torch.float32)
    ld, yd, td = dev(lengths), dev(labels), dev(tlens)
    dmd = dev(dm, torch.float32) if drop else None
    logp = s2s.attention_forward(gctx, cfg, Pd, hd, yd, lengths=ld, tlens=td, dropmask=dmd, lam=lam)
    alpha = s2s.attention_get(gctx, s2s.GET_ALPHA, (B, T, L)).cpu().numpy()
    pen = s2s.attention_get(gctx, s2s.GET_PENALTY, (B, T)).cpu().numpy()
    G = torch.zeros_like(Pd)
    dh = s2s.attention_backward(gctx, cfg, Pd, G, hd, yd, dev(dlogp, torch.float32), lengths=ld, tlens=td, dropmask=dmd, lam=lam)
    logp = logp.cpu().numpy(); dh = dh.cpu().numpy(); G = G.cpu().numpy()
    Gref = np.zeros_like(P)
    for b in range(B):
        Lb, Tb = lengths[b], tlens[b]
        dmb = dm[b, :Tb] if drop else None
        ref = orc64.attention_forward(cfg, P, h[b, :Lb], labels[b, :Tb], lam=lam, dropmask=dmb)
        assert rel_err(logp[b, :Tb], ref["logp"]) < TOL
        assert rel_err(alpha[b, :Tb, :Lb], ref["alpha"]) < TOL
        assert np.abs(pen[b, :Tb] - ref["pen"]).max() < TOL * max(1.0, np.abs(ref["pen"]).max())
        Gb, dhb = orc64.attention_backward(cfg, P, h[b, :Lb], labels[b, :Tb], dlogp[b, :Tb], lam=lam, dropmask=dmb)
        Gref += Gb
        assert rel_err(dh[b, :Lb], dhb) < TOL
        assert np.abs(dh[b, Lb:]).max() == 0.0 if Lb < L else True
    errs = _seg_errs(cfg, orc64, G, Gref)
    bad = {k: v for k, v in errs.items() if v > TOL}
    assert not bad, bad


@pytest.mark.parametrize("K,KF", [(0, 4), (3, 6)])
def test_model_fwdbwd_matches_oracle(s2s, gctx, orc64, K, KF):
    cfg = dict(MID, K=K, KF=KF)
    B, L, T = 5, 40, 6
    P = init_params(cfg, seed=11, dtype=np.float64, oracle=orc64) * 1.5
    X, lengths, labels, tlens = make_batch(cfg, B, L, T, seed=3)
    ref = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, normalize_nll=True, nthreads=4)
    Pd = dev(P, torch.float32); G = torch.zeros_like(Pd)
    Xd, ld, yd, td = dev(X), dev(lengths), dev(labels), dev(tlens)
    logp = gctx.new(B, T, cfg["V"]); dX = gctx.new(B, L, cfg["D"])
    nll = s2s.model_fwdbwd(gctx, cfg, Pd, G, Xd, yd, lengths=ld, tlens=td, flags=s2s.NORMALIZE_NLL, logp=logp, dX=dX)
    annot = s2s.model_annotations(gctx, B, L, 2 * cfg["H"]).cpu().numpy()
    assert rel_err(nll.cpu().numpy(), ref["nll"]) < TOL
    logp = logp.cpu().numpy(); dX = dX.cpu().numpy()
    for b in range(B):
        Lb, Tb = lengths[b], tlens[b]
        assert rel_err(annot[b, :Lb], ref["annot"][b, :Lb]) < TOL
        assert rel_err(logp[b, :Tb], ref["logp"][b, :Tb]) < TOL
        assert rel_err(dX[b, :Lb], ref["dX"][b, :Lb]) < TOL
    errs = _seg_errs(cfg, orc64, G.cpu().numpy(), ref["G"])
    bad = {k: v for k, v in errs.items() if v > TOL}
    assert not bad, bad
    # forward-only entry gives the same loss
    nll2, _ = s2s.model_forward(gctx, cfg, Pd, Xd, yd, lengths=ld, tlens=td, flags=s2s.NORMALIZE_NLL)
    assert torch.equal(nll2, nll)


def test_batch_equals_per_utterance(s2s, gctx, orc64):
    # invariant lifted from the reference notebooks (Attention.ipynb:1202/1763): a batch gives the
    # same rows as utterance-at-a-time ("SGD mode") calls
    cfg = dict(MID)
    B, L, T = 3, 33, 5
    P = dev(init_params(cfg, seed=2, dtype=np.float64, oracle=orc64), torch.float32)
    X, lengths, labels, tlens = make_batch(cfg, B, L, T, seed=8)
    nll_b, logp_b = s2s.model_forward(gctx, cfg, P, dev(X), dev(labels), lengths=dev(lengths), tlens=dev(tlens))
    for b in range(B):
        Lb, Tb = int(lengths[b]), int(tlens[b])
        nll_1, logp_1 = s2s.model_forward(gctx, cfg, P, dev(X[b:b + 1, :Lb]), dev(labels[b:b + 1, :Tb]))
        assert rel_err(logp_1[0].cpu().numpy(), logp_b[b, :Tb].cpu().numpy()) < 1e-5
        assert abs(float(nll_1[0]) - float(nll_b[b])) < 1e-4 * abs(float(nll_b[b]))


def test_optimizer_and_noise_match_oracle(s2s, gctx, orc32, orc64):
    rng = np.random.default_rng(0)
    n = 10007
    w = rng.standard_normal(n).astype(np.float32) * 0.1; eps = rng.standard_normal(n).astype(np.float32)
    assert rel_err(s2s.weightnoise_sample(gctx, dev(w), 0.075, eps=dev(eps)).cpu().numpy(), orc64.weightnoise_sample(w, eps, 0.075)) < 1e-6
    weight = np.concatenate([w, np.log(0.075 ** 2) + 0.1 * rng.standard_normal(n)]).astype(np.float32)
    assert rel_err(s2s.awn_sample(gctx, dev(weight), eps=dev(eps)).cpu().numpy(), orc64.awn_sample(weight, eps)) < 1e-5
    Lg = s2s.awn_forward(gctx, dev(weight), 1.0, 3.25); Lr = orc64.awn_forward(weight, 1.0, 3.25)
    assert abs(Lg - Lr) < 1e-5 * abs(Lr)
    g = rng.standard_normal(n).astype(np.float32)
    assert rel_err(s2s.awn_accgrad(gctx, dev(weight), dev(g), 1.0).cpu().numpy(), orc64.awn_accgrad(weight, g, 1.0)) < 1e-5
    # gradient step: /B, clip, weight decay, injected noise, adadelta, row-norm constraint
    gd = dev(g.copy()); noise = rng.standard_normal(n).astype(np.float32)
    nrm = s2s.grad_finalize(gctx, gd, dev(w), 8, 0.5, wd=1e-3, noise=dev(noise), noise_sigma=0.01)
    gr = g.astype(np.float64).copy()
    nr = orc64.grad_finalize(gr, w, 8, 0.5, 1e-3, noise, 0.01)
    assert abs(nrm - nr) < 1e-5 * nr and rel_err(gd.cpu().numpy(), gr) < 1e-5
    x, v, a = dev(w.copy()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    xr, vr, ar = w.astype(np.float64).copy(), np.zeros(n), np.zeros(n)
    for _ in range(3):
        s2s.adadelta(gctx, x, gd, v, a); orc64.adadelta(xr, gr, vr, ar)
    assert rel_err(x.cpu().numpy(), xr) < 1e-5
    Wm = (rng.standard_normal((37, 53)) * rng.uniform(0.01, 0.4, (37, 1))).astype(np.float32)
    Wd = dev(Wm.copy()); Wr = Wm.astype(np.float64).copy()
    assert s2s.rownorm_constraint(gctx, Wd, 1.0) == 0 and orc64.rownorm_constraint(Wr, 1.0) == 0
    assert rel_err(Wd.cpu().numpy(), Wr) < 1e-6
    Wn = Wm.copy(); Wn[3, 4] = np.nan
    assert s2s.rownorm_constraint(gctx, dev(Wn), 1.0) == 1
    # Philox sampler: right moments
    z = (s2s.weightnoise_sample(gctx, torch.zeros(1 << 20, device="cuda"), 1.0, seed=7)).cpu().numpy()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3


@pytest.mark.parametrize("extra", [{}, dict(MLP=2, NL=0)])
def test_beam_search_matches_oracle(s2s, gctx, orc32, orc64, extra):
    cfg = dict(MID, K=3, KF=4, **extra)
    P = init_params(cfg, seed=21, dtype=np.float64, oracle=orc64) * 3.0
    rng = np.random.default_rng(4)
    for trial in range(3):
        L = 20 + 7 * trial
        h = rng.standard_normal((L, 2 * cfg["H"]))
        ref_y, ref_lp = orc64.beam_search(cfg, P, h, eos=cfg["V"] - 1, K=4, maxlen=12)
        y, lp = s2s.beam_search(gctx, cfg, dev(P, torch.float32), dev(h, torch.float32), eos=cfg["V"] - 1, beam=4, maxlen=12)
        assert list(ref_y) == y            # bit-exact label sequence
        assert abs(lp - ref_lp) < 1e-3 * max(1.0, abs(ref_lp))


@pytest.mark.parametrize("beam,maxlen", [(1, 12), (4, 3), (5, 20), (8, 40), (16, 25)])
def test_beam_search_device_bookkeeping_equals_host_and_oracle(s2s, gctx, orc64, beam, maxlen):
    """The device-resident search (top-k, finished list, label sequences and beam scores in HBM; the host reads 16 bytes of
    status every 8 labels) against the first version that copied the scores to the host every label (S2S_BEAM_HOST=1) and
    against the oracle's Attention:BeamSearch (Attention.lua:332-438): identical label sequences, equal scores.  Covers
    the greedy case, termination by the length limit (maxlen 3), beams wider than the surviving hypotheses and maxlen not a
    multiple of the status interval."""
    import os
    cfg = dict(MID, K=3, KF=4)
    P = init_params(cfg, seed=33, dtype=np.float64, oracle=orc64) * 3.0
    rng = np.random.default_rng(beam * 100 + maxlen)
    eos = cfg["V"] - 1
    old = os.environ.get("S2S_BEAM_HOST")
    try:
        for trial in range(2):
            L = 17 + 9 * trial
            h = rng.standard_normal((L, 2 * cfg["H"]))
            Pd, hd = dev(P, torch.float32), dev(h, torch.float32)
            os.environ["S2S_BEAM_HOST"] = "0"
            y_dev, lp_dev = s2s.beam_search(gctx, cfg, Pd, hd, eos=eos, beam=beam, maxlen=maxlen)
            os.environ["S2S_BEAM_HOST"] = "1"
            y_host, lp_host = s2s.beam_search(gctx, cfg, Pd, hd, eos=eos, beam=beam, maxlen=maxlen)
            assert y_dev == y_host and lp_dev == lp_host          # same kernels, same fp32 adds: bit-identical
            ref_y, ref_lp = orc64.beam_search(cfg, P, h, eos=eos, K=beam, maxlen=maxlen)
            assert list(ref_y) == y_dev
            assert abs(lp_dev - ref_lp) < 1e-3 * max(1.0, abs(ref_lp))
    finally:
        if old is None:
            os.environ.pop("S2S_BEAM_HOST", None)
        else:
            os.environ["S2S_BEAM_HOST"] = old


def test_caller_defined_graph_replays_the_same_results(s2s, gctx, orc64):
    # s2s_graph_begin / _end / _launch: a captured sequence of library calls (decoder forward + loss seed + backward)
    # replayed on new input values must give what the eager calls give
    cfg = dict(MID, NL=0, MLP=2)
    B, L, T = 3, 40, 6
    P = dev(init_params(cfg, seed=9, dtype=np.float64, oracle=orc64) * 1.5, torch.float32)
    rng = np.random.default_rng(8)
    h = dev(rng.standard_normal((B, L, 2 * cfg["H"])), torch.float32)
    labels = dev(rng.integers(0, cfg["V"] - 1, (B, T)).astype(np.int32))
    G = torch.zeros_like(P); nll = torch.zeros(B, device="cuda"); dlogp = torch.zeros(B, T, cfg["V"], device="cuda")

    def run():
        logp = s2s.attention_forward(gctx, cfg, P, h, labels)
        s2s.nll_grad_seed(gctx, logp, labels, nll=nll, dlogp=dlogp)
        return logp, s2s.attention_backward(gctx, cfg, P, G, h, labels, dlogp)

    run()                                       # eager warm-up sizes the workspaces
    gctx.graph_begin()
    logp_g, dh_g = run()
    gid = gctx.graph_end()
    try:
        h.copy_(dev(rng.standard_normal((B, L, 2 * cfg["H"])), torch.float32))     # new values, same buffers
        G.zero_()
        gctx.graph_launch(gid)
        torch.cuda.synchronize()
        got = (logp_g.clone(), dh_g.clone(), G.clone(), nll.clone())
    finally:
        gctx.graph_destroy(gid)
    G.zero_()
    logp_e, dh_e = run()
    torch.cuda.synchronize()
    for a, b in zip(got, (logp_e, dh_e, G, nll)):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
