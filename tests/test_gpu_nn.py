"""The reference-facing module surface (nn.Attention, nn.RNN(nn.GRU), nn.TemporalConvolutionZeroBias, ...)
through forward/backward, in single-utterance ("SGD") and batch mode, against the oracle."""
import numpy as np
import pytest
import torch

from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_rnn_gru_module_sgd_and_batch(s2s, gctx, orc64):
    nn = s2s.nn
    rng = np.random.default_rng(0)
    gru = nn.GRU(gctx, 24, 128)
    W = gru.weight.cpu().numpy().astype(np.float64)
    for reverse in (False, True):
        rnn = nn.RNN(gctx, gru, reverse)
        x = rng.standard_normal((31, 24))
        y_ref, gates = orc64.gru_seq_forward(W[0], W[1], W[2], x, reverse=reverse)
        y = rnn.forward(dev(x, torch.float32))                       # 2-D = one utterance
        assert y.shape == (31, 128) and rel_err(y.cpu().numpy(), y_ref) < TOL
        dy = rng.standard_normal((31, 128))
        gru.zeroGradParameters()
        dx = rnn.backward(dev(x, torch.float32), dev(dy, torch.float32))
        dx_ref, dWz, dWr, dWh = orc64.gru_seq_backward(W[0], W[1], W[2], x, y_ref, gates, dy, reverse=reverse)
        assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
        assert rel_err(gru.gradWeight.cpu().numpy(), np.stack([dWz, dWr, dWh])) < TOL
        xb = np.stack([x, x[::-1].copy()])                           # 3-D = batch; rows equal the SGD-mode result
        yb = rnn.forward(dev(xb, torch.float32))
        assert rel_err(yb[0].cpu().numpy(), y_ref) < TOL


def test_attention_module_sgd_equals_batch(s2s, gctx, orc64):
    # the reference notebook's invariant (Attention.ipynb:1202/1763): batch rows == SGD-mode result
    nn = s2s.nn
    rng = np.random.default_rng(1)
    # built the way timit/model_chorowski_baseline.lua:48-71 builds it: a GRU(st, st) recurrent part and a Maxout-Linear-LogSoftMax MLP
    decoder_recurrent = nn.Sequential(gctx, nn.GRU(gctx, 64, 64))
    decoder_mlp = nn.Sequential(gctx, nn.Maxout(gctx, 64 + 256, 8, 3), nn.Linear(gctx, 8, 11), nn.LogSoftMax(gctx))
    att = nn.Attention(gctx, decoder_recurrent, decoder_mlp, 128, 4, 3, 64, 256, 11, False, 0.0)
    assert att.cfg["M"] == 8 and att.cfg["MW"] == 3 and att.cfg["MLP"] == 1 and att.cfg["NL"] == 0
    # the sub-modules' initial values are adopted (a pre-loaded sub-module keeps its weights)
    assert torch.equal(att._p["Gz"], decoder_recurrent.modules[0].weight[0]) and torch.equal(att._p["Wo"], decoder_mlp.modules[1].weight)
    L, T, V = 37, 5, 11
    h = rng.standard_normal((2, L, 256)) * 0.5
    lab = rng.integers(0, V, (2, T))
    y = np.eye(V)[lab]
    out_b = att.forward([dev(h, torch.float32), dev(y, torch.float32)])
    alpha_b = att.alpha()
    P = att.flat.cpu().numpy().astype(np.float64)
    for b in range(2):
        out_s = att.forward([dev(h[b], torch.float32), dev(y[b], torch.float32)])
        assert out_s.shape == (T, V)
        assert rel_err(out_s.cpu().numpy(), out_b[b].cpu().numpy()) < 1e-5
        ref = orc64.attention_forward(att.cfg, P, h[b], lab[b])
        assert rel_err(out_s.cpu().numpy(), ref["logp"]) < TOL
        assert rel_err(att.alpha().cpu().numpy(), ref["alpha"]) < TOL
        assert rel_err(alpha_b[b].cpu().numpy(), ref["alpha"]) < TOL
    # backward accumulates parameter gradients inside updateGradInput (Attention.lua:325)
    att.zeroGradParameters()
    dlogp = rng.standard_normal((T, V))
    att.forward([dev(h[0], torch.float32), dev(y[0], torch.float32)])
    dh, _ = att.backward([dev(h[0], torch.float32), dev(y[0], torch.float32)], dev(dlogp, torch.float32))
    G_ref, dh_ref = orc64.attention_backward(att.cfg, P, h[0], lab[0], dlogp)
    assert rel_err(dh.cpu().numpy(), dh_ref) < TOL
    assert rel_err(att.gradFlat.cpu().numpy(), G_ref) < TOL
    with pytest.raises(s2s.S2SError):
        att.forward([dev(h[0, 0], torch.float32), dev(y[0], torch.float32)])      # "x must be 2d or 3d"


def test_tconv_zero_bias_module(s2s, gctx):
    nn = s2s.nn
    m = nn.TemporalConvolutionZeroBias(gctx, 5, 4, 1)
    m.weight.copy_(torch.arange(1, 21, dtype=torch.float32).view(4, 5))
    y = m.forward(torch.ones(10, 5, device="cuda"))
    assert np.array_equal(y.cpu().numpy(), np.tile([15, 40, 65, 90], (10, 1)).astype(np.float32))   # Attention.ipynb:123-154
    m.backward(torch.ones(10, 5, device="cuda"), torch.ones(10, 4, device="cuda"))
    assert float(m.gradBias.abs().sum()) == 0.0 and float(m.bias.abs().sum()) == 0.0
    assert np.allclose(m.gradWeight.cpu().numpy(), 10.0)
    with pytest.raises(s2s.S2SError):
        nn.TemporalConvolutionZeroBias(gctx, 5, 4, 3)


def test_gru_single_step_module(s2s, gctx, orc64):
    # nn.GRU stepped by hand with an explicit previous state (GRU.lua:22-38 through nn.Recurrent)
    nn = s2s.nn
    rng = np.random.default_rng(3)
    Din, H, B = 123, 64, 5
    gru = nn.GRU(gctx, Din, H)
    W = gru.weight.cpu().numpy().astype(np.float64)
    x = rng.standard_normal((B, Din)); hp = rng.standard_normal((B, H)); dhn = rng.standard_normal((B, H))
    hn = gru.forward([dev(x, torch.float32), dev(hp, torch.float32)])
    dx, dhp = gru.backward([dev(x, torch.float32), dev(hp, torch.float32)], dev(dhn, torch.float32))
    dW_ref = np.zeros_like(W)
    for b in range(B):
        h_ref, z, r, hc = orc64.gru_step_forward(W[0], W[1], W[2], x[b], hp[b])
        assert rel_err(hn[b].cpu().numpy(), h_ref) < TOL
        dxb, dhpb, dWz, dWr, dWh = orc64.gru_step_backward(W[0], W[1], W[2], x[b], hp[b], z, r, hc, dhn[b])
        assert rel_err(dx[b].cpu().numpy(), dxb) < TOL and rel_err(dhp[b].cpu().numpy(), dhpb) < TOL
        dW_ref += np.stack([dWz, dWr, dWh])
    assert rel_err(gru.gradWeight.cpu().numpy(), dW_ref) < TOL
    # zero initial state when prev_h is omitted (Recurrent.lua:110-112)
    h0 = gru.forward(dev(x[0], torch.float32))
    h0_ref, *_ = orc64.gru_step_forward(W[0], W[1], W[2], x[0], np.zeros(H))
    assert rel_err(h0.cpu().numpy(), h0_ref) < TOL


def test_dropout_mask_statistics(s2s, gctx):
    m = s2s.dropout_mask(gctx, (1 << 20,), 0.5, seed=11).cpu().numpy()
    vals = np.unique(m)
    assert set(vals.tolist()) == {0.0, 2.0}
    assert abs((m == 0).mean() - 0.5) < 5e-3
    m2 = s2s.dropout_mask(gctx, (1 << 20,), 0.2, seed=11).cpu().numpy()
    assert abs((m2 == 0).mean() - 0.2) < 5e-3 and abs(m2.max() - 1.25) < 1e-6


@pytest.mark.parametrize("peep", [False, True])
def test_lstm_single_step_module(s2s, gctx, peep):
    # nn.LSTM stepped by hand with explicit {prev_h, prev_c} (LSTM.lua:100-136) against the float64 autograd restatement
    from tests import torch_ref
    nn = s2s.nn
    rng = np.random.default_rng(11 + int(peep))
    Din, H, B = 36, 64, 5
    lstm = nn.LSTM(gctx, Din, H, peepholes=peep)
    P = lstm.weight.cpu().double()
    x = rng.standard_normal((B, Din)); hp = rng.standard_normal((B, H)) * 0.5; cp = rng.standard_normal((B, H)) * 0.5
    dhn = rng.standard_normal((B, H)); dcn = rng.standard_normal((B, H))
    f32 = lambda a: dev(a, torch.float32)
    hn, cn = lstm.forward([f32(x), f32(hp), f32(cp)])
    dx, dhp, dcp = lstm.backward([f32(x), f32(hp), f32(cp)], [f32(dhn), f32(dcn)])
    Pt = P.clone().requires_grad_(True)
    xt, hpt, cpt = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, hp, cp))
    gates = torch_ref.lstm_unpack(Pt, Din, H, peep)
    outs = [torch_ref.lstm_step(gates, xt[b], hpt[b], cpt[b]) for b in range(B)]
    hr = torch.stack([o[0] for o in outs]); cr = torch.stack([o[1] for o in outs])
    ((hr * torch.tensor(dhn)).sum() + (cr * torch.tensor(dcn)).sum()).backward()
    assert rel_err(hn.cpu().numpy(), hr.detach().numpy()) < TOL and rel_err(cn.cpu().numpy(), cr.detach().numpy()) < TOL
    assert rel_err(dx.cpu().numpy(), xt.grad.numpy()) < TOL
    assert rel_err(dhp.cpu().numpy(), hpt.grad.numpy()) < TOL and rel_err(dcp.cpu().numpy(), cpt.grad.numpy()) < TOL
    assert rel_err(lstm.gradWeight.cpu().numpy(), Pt.grad.numpy()) < TOL
    # zero initial state when prev_h / prev_c are omitted (LSTM.lua:108-109)
    h0, c0 = lstm.forward([f32(x[0])])
    h0r, c0r = torch_ref.lstm_step(torch_ref.lstm_unpack(P, Din, H, peep), torch.tensor(x[0]), torch.zeros(H, dtype=torch.float64), torch.zeros(H, dtype=torch.float64))
    assert rel_err(h0.cpu().numpy(), h0r.numpy()) < TOL and rel_err(c0.cpu().numpy(), c0r.numpy()) < TOL
