"""Cross-checks the oracle's hand-written forward/backward (oracle/s2s_oracle.c) against an
independent float64 autograd derivation (tests/torch_ref.py), and the invariants lifted from the
reference notebooks: custom NLL == ClassNLL, batch == per-utterance."""
import numpy as np
import pytest
import torch

from oracle.oracle import init_params
from tests import torch_ref

SMALL = dict(D=5, H=4, NL=2, S=6, ST=5, V=7, K=0, KF=4, M=3, MW=2)


def _data(cfg, B, Lmax, Tmax, seed, ragged=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, Lmax, cfg["D"]))
    lengths = rng.integers(Lmax // 2, Lmax + 1, B).astype(np.int32) if ragged else np.full(B, Lmax, np.int32)
    tlens = rng.integers(2, Tmax + 1, B).astype(np.int32) if ragged else np.full(B, Tmax, np.int32)
    lengths[0] = Lmax; tlens[0] = Tmax
    labels = rng.integers(0, cfg["V"] - 1, (B, Tmax)).astype(np.int32)
    for b in range(B):
        labels[b, tlens[b] - 1] = cfg["V"] - 1  # EOS last
    return X, lengths, labels, tlens


@pytest.mark.parametrize("K,KF,lam", [(0, 4, 0.0), (3, 4, 0.0), (3, 5, 0.0), (2, 10, 0.01), (0, 4, 0.05)])
def test_model_matches_autograd_f64(orc64, K, KF, lam):
    cfg = dict(SMALL, K=K, KF=KF)
    P = init_params(cfg, seed=7, dtype=np.float64, oracle=orc64) * 2.0
    X, lengths, labels, tlens = _data(cfg, 3, 9, 5, seed=11)
    ref = torch_ref.model_fwdbwd(cfg, P, X, lengths, labels, tlens, lam=lam)
    out = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, lam=lam)
    assert np.allclose(out["nll"], ref["nll"], rtol=1e-10, atol=1e-12)
    for b in range(3):
        L, T = lengths[b], tlens[b]
        assert np.allclose(out["logp"][b, :T], ref["logp"][b], rtol=1e-9, atol=1e-11)
        assert np.allclose(out["alpha"][b, :T, :L], ref["alpha"][b], rtol=1e-9, atol=1e-12)
        assert np.allclose(out["annot"][b, :L], ref["annot"][b], rtol=1e-10, atol=1e-12)
        assert np.allclose(out["dX"][b, :L], ref["dX"][b], rtol=1e-7, atol=1e-11)
    g, gr = out["G"], ref["G"]
    assert np.abs(g - gr).max() <= 1e-9 * max(1.0, np.abs(gr).max())
    # every parameter segment must receive a gradient somewhere (except dead biases)
    assert np.count_nonzero(gr) > 0.9 * (gr.size - 3 * cfg["S"])


def test_model_vgg_decoder_mlp_matches_autograd_f64(orc64):
    # MLP = 2: Maxout-Linear-Maxout-Linear-LogSoftMax of librispeech/model_vgg.lua:76-80
    cfg = dict(SMALL, MLP=2, M=4, MW=3)
    P = init_params(cfg, seed=3, dtype=np.float64, oracle=orc64) * 2.0
    assert P.size == orc64.param_count(dict(cfg, MLP=1)) + cfg["M"] * cfg["M"] + cfg["M"] + cfg["M"] * cfg["MW"] * (cfg["M"] + 1)
    X, lengths, labels, tlens = _data(cfg, 3, 8, 5, seed=5)
    ref = torch_ref.model_fwdbwd(cfg, P, X, lengths, labels, tlens)
    out = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens)
    assert np.allclose(out["nll"], ref["nll"], rtol=1e-10, atol=1e-12)
    for b in range(3):
        assert np.allclose(out["logp"][b, :tlens[b]], ref["logp"][b], rtol=1e-9, atol=1e-11)
    g, gr = out["G"], ref["G"]
    assert np.abs(g - gr).max() <= 1e-9 * max(1.0, np.abs(gr).max())
    segs = dict(zip(torch_ref.segment_names(cfg), orc64.param_segments(cfg)))
    for name in ("Wl", "bl", "Wm2", "bm2"):
        off, rows, cols = segs[name]
        assert np.count_nonzero(gr[off:off + rows * cols]) > 0, name


def test_model_flags_and_dropout(orc64):
    cfg = dict(SMALL, K=2, KF=3)
    P = init_params(cfg, seed=3, dtype=np.float64, oracle=orc64)
    X, lengths, labels, tlens = _data(cfg, 2, 8, 4, seed=5)
    rng = np.random.default_rng(0)
    dm = (rng.random((2, 4, cfg["ST"] + 2 * cfg["H"])) > 0.5) / 0.5
    ref = torch_ref.model_fwdbwd(cfg, P, X, lengths, labels, tlens, dropmask=dm, normalize_nll=True, normalize_grad=True)
    out = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, dropmask=dm, normalize_nll=True, normalize_grad=True)
    assert np.allclose(out["nll"], ref["nll"], rtol=1e-10)
    assert np.abs(out["G"] - ref["G"]).max() <= 1e-9 * max(1.0, np.abs(ref["G"]).max())


def test_f32_oracle_close_to_f64(orc32, orc64):
    cfg = dict(SMALL, K=3, KF=4)
    P = init_params(cfg, seed=9, dtype=np.float64, oracle=orc64)
    X, lengths, labels, tlens = _data(cfg, 4, 12, 6, seed=2)
    a = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, nthreads=1)
    b = orc32.model_fwdbwd(cfg, P, X, lengths, labels, tlens, nthreads=4)   # also exercises the OpenMP reduction
    assert np.allclose(a["nll"], b["nll"], rtol=1e-5)
    assert np.abs(a["G"] - b["G"]).max() <= 1e-4 * np.abs(a["G"]).max()


def test_nll_equals_classnll(orc64):
    # AttentionSmallModel.ipynb:304-305,350-351: -sum(labelmask*logp) == summed ClassNLLCriterion
    cfg = SMALL
    P = init_params(cfg, seed=1, dtype=np.float64, oracle=orc64)
    X, lengths, labels, tlens = _data(cfg, 2, 7, 4, seed=8, ragged=False)
    out = orc64.model_fwdbwd(cfg, P, X, lengths, labels, tlens, backward=False)
    for b in range(2):
        lp = torch.tensor(out["logp"][b])
        nll = torch.nn.functional.nll_loss(lp, torch.tensor(labels[b]).long(), reduction="sum")
        assert abs(float(nll) - out["nll"][b]) < 1e-12


def test_gru_step_and_seq(orc64):
    rng = np.random.default_rng(4)
    D, H, L = 6, 5, 7
    Wz, Wr, Wh = (rng.standard_normal((H, H + D)) * 0.4 for _ in range(3))
    x = rng.standard_normal((L, D)); dy = rng.standard_normal((L, H))
    for rev in (False, True):
        y, gates = orc64.gru_seq_forward(Wz, Wr, Wh, x, rev)
        tw = [torch.tensor(w, requires_grad=True) for w in (Wz, Wr, Wh)]
        tx = torch.tensor(x, requires_grad=True)
        ty = torch_ref.gru_seq(*tw, tx, rev)
        assert np.allclose(y, ty.detach().numpy(), rtol=1e-12)
        ty.backward(torch.tensor(dy))
        dx, dWz, dWr, dWh = orc64.gru_seq_backward(Wz, Wr, Wh, x, y, gates, dy, rev)
        assert np.allclose(dx, tx.grad.numpy(), rtol=1e-9, atol=1e-12)
        for a, b in zip((dWz, dWr, dWh), tw):
            assert np.allclose(a, b.grad.numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("peep", [False, True])
def test_lstm_seq(orc64, peep):
    rng = np.random.default_rng(6)
    D, H, L = 4, 3, 6
    n = orc64.lstm_param_count(D, H, peep)
    P = rng.standard_normal(n) * 0.5
    x = rng.standard_normal((L, D)); dy = rng.standard_normal((L, H))
    for rev in (False, True):
        y, c, acts = orc64.lstm_seq_forward(P, D, H, peep, x, rev)
        tP = torch.tensor(P, requires_grad=True); tx = torch.tensor(x, requires_grad=True)
        ty = torch_ref.lstm_seq(tP, D, H, peep, tx, rev)
        assert np.allclose(y, ty.detach().numpy(), rtol=1e-12)
        ty.backward(torch.tensor(dy))
        dx, dP = orc64.lstm_seq_backward(P, D, H, peep, x, y, c, acts, dy, rev)
        assert np.allclose(dx, tx.grad.numpy(), rtol=1e-9, atol=1e-12)
        assert np.allclose(dP, tP.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_attention_standalone_and_introspection(orc64):
    cfg = dict(SMALL, K=2, KF=4)
    P = init_params(cfg, seed=12, dtype=np.float64, oracle=orc64)
    rng = np.random.default_rng(13)
    L, T = 9, 5
    h = rng.standard_normal((L, 2 * cfg["H"])); labels = rng.integers(0, cfg["V"], T).astype(np.int32)
    out = orc64.attention_forward(cfg, P, h, labels, lam=0.02)
    Pt = torch.tensor(P, requires_grad=True); ht = torch.tensor(h, requires_grad=True)
    logp, alpha, aux = torch_ref.decoder(cfg, torch_ref.unflatten(cfg, Pt), ht, [int(v) for v in labels], lam=0.02)
    assert np.allclose(out["logp"], logp.detach().numpy(), rtol=1e-10)
    assert np.allclose(out["alpha"], alpha.detach().numpy(), rtol=1e-10)
    assert np.allclose(out["alpha"].sum(1), 1.0)
    assert np.allclose(out["q"], aux["q"].detach().numpy(), rtol=1e-10)
    assert np.allclose(out["Vh"], aux["Vh"].detach().numpy(), rtol=1e-10)
    dlogp = rng.standard_normal((T, cfg["V"]))
    logp.backward(torch.tensor(dlogp))
    G, dh = orc64.attention_backward(cfg, P, h, labels, dlogp, lam=0.02)
    assert np.allclose(dh, ht.grad.numpy(), rtol=1e-8, atol=1e-11)
    assert np.abs(G - Pt.grad.numpy()).max() < 1e-9 * max(1, np.abs(G).max())


def test_beam_search_greedy_consistency(orc64):
    # K=1 beam == greedy argmax decode fed back (Attention.lua:366-437 with K=1)
    cfg = dict(SMALL, K=2, KF=3)
    P = init_params(cfg, seed=21, dtype=np.float64, oracle=orc64) * 3
    rng = np.random.default_rng(22)
    L = 8
    h = rng.standard_normal((L, 2 * cfg["H"]))
    eos = cfg["V"] - 1
    seq, lp = orc64.beam_search(cfg, P, h, eos, K=1, maxlen=6)
    # greedy by repeated teacher-forced forward
    lab = []
    for t in range(7):
        out = orc64.attention_forward(cfg, P, h, np.array(lab + [0], dtype=np.int32))
        nxt = int(out["logp"][t].argmax()); lab.append(nxt)
        if nxt == eos or len(lab) == 7:
            break
    assert list(seq) == lab[:len(seq)]
    seq5, lp5 = orc64.beam_search(cfg, P, h, eos, K=5, maxlen=6)
    assert lp5 >= lp - 1e-12 and 1 <= len(seq5) <= 7
