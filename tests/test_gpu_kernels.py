"""GPU parity of the individual kernels, called through the C ABI (ctypes), against float64 numpy /
the CPU oracle.  Tolerance: north_star's fp32 bound, <= 1e-4 relative (max-abs error over the
reference's max-abs value)."""
import numpy as np
import pytest
import torch

from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(9, 7, 5), (130, 129, 123), (257, 64, 379), (1000, 300, 64)])
def test_gemm_simt(s2s, gctx, tA, tB, M, N, K):
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = 0.5 * ((A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)) + 2.0 * C0 + bias
    Cd = dev(C0)
    s2s.gemm(gctx, dev(A), dev(B), tA=tA, tB=tB, alpha=0.5, beta=2.0, C_out=Cd, bias=dev(bias), impl=1)
    assert rel_err(Cd.cpu().numpy(), ref) < 1e-5


def _attn_ref(Vh, h, q, w, lengths):
    B, L, S = Vh.shape
    alpha = np.zeros((B, L)); c = np.zeros((B, h.shape[2]))
    for b in range(B):
        Lb = lengths[b]
        e = np.tanh(q[b][None, :] + Vh[b, :Lb]) @ w
        p = np.exp(e - e.max()); p /= p.sum()
        alpha[b, :Lb] = p
        c[b] = p @ h[b, :Lb]
    return alpha, c


@pytest.mark.parametrize("B,L,S,A", [(3, 70, 128, 128), (2, 33, 256, 512), (4, 300, 512, 512), (1, 1, 128, 256), (5, 97, 512, 256)])
def test_attn_step_forward_backward(s2s, gctx, B, L, S, A):
    rng = np.random.default_rng(B * 100 + L)
    Vh = rng.standard_normal((B, L, S)); h = rng.standard_normal((B, L, A)); q = rng.standard_normal((B, S))
    w = rng.standard_normal(S) / np.sqrt(S)
    lengths = rng.integers(max(1, L // 3), L + 1, B).astype(np.int32); lengths[0] = L
    alpha_ref, c_ref = _attn_ref(Vh, h, q, w, lengths)
    dVh, dh, dq_, dw = [dev(x, torch.float32) for x in (Vh, h, q, w)]
    dl = dev(lengths)
    alpha, c = s2s.attn_step_forward(gctx, dVh, dh, dq_, dw, lengths=dl)
    assert rel_err(alpha.cpu().numpy(), alpha_ref) < TOL
    assert rel_err(c.cpu().numpy(), c_ref) < TOL
    # padded tail is exactly zero
    for b in range(B):
        assert float(alpha[b, lengths[b]:].abs().sum()) == 0.0

    # backward: dq, de given dc and an incoming dalpha
    dc = rng.standard_normal((B, A)); dain = rng.standard_normal((B, L)) * 0.1
    dq_ref = np.zeros((B, S)); de_ref = np.zeros((B, L))
    for b in range(B):
        Lb = lengths[b]
        a = alpha_ref[b, :Lb]
        da = h[b, :Lb] @ dc[b] + dain[b, :Lb]
        de = a * (da - a @ da)
        th = np.tanh(q[b][None, :] + Vh[b, :Lb])
        dq_ref[b] = (de[:, None] * w[None, :] * (1 - th * th)).sum(0)
        de_ref[b, :Lb] = de
    dq_g, de_g = s2s.attn_step_backward(gctx, dVh, dh, dq_, dw, dev(alpha_ref, torch.float32), dev(dc, torch.float32),
                                        dalpha_in=dev(dain, torch.float32), lengths=dl)
    assert rel_err(de_g.cpu().numpy(), de_ref) < TOL
    assert rel_err(dq_g.cpu().numpy(), dq_ref, floor=1e-2) < TOL   # L = 1 gives de = 0 exactly: absolute floor


@pytest.mark.parametrize("H,Din,B,L,ndir,reverse", [(128, 20, 5, 37, 1, False), (128, 20, 5, 37, 1, True), (256, 123, 6, 50, 2, False),
                                                    (128, 256, 3, 9, 2, False), (256, 512, 9, 21, 2, False)])
def test_gru_seq_matches_oracle(s2s, gctx, orc64, H, Din, B, L, ndir, reverse):
    rng = np.random.default_rng(H + Din + B)
    W = (rng.uniform(-1, 1, (ndir * 3, H, H + Din)) / np.sqrt(H + Din) * 1.5)
    x = rng.standard_normal((B, L, Din))
    lengths = rng.integers(max(1, L // 2), L + 1, B).astype(np.int32); lengths[0] = L
    dy = rng.standard_normal((B, L, ndir * H))
    for b in range(B):
        x[b, lengths[b]:] = 0; dy[b, lengths[b]:] = 0
    y_ref = np.zeros((B, L, ndir * H)); dx_ref = np.zeros_like(x); dW_ref = np.zeros_like(W)
    for b in range(B):
        Lb = lengths[b]
        for d in range(ndir):
            rv = (d == 1) if ndir == 2 else reverse
            Wz, Wr, Wh = W[3 * d], W[3 * d + 1], W[3 * d + 2]
            yb, gates = orc64.gru_seq_forward(Wz, Wr, Wh, x[b, :Lb], reverse=rv)
            y_ref[b, :Lb, d * H:(d + 1) * H] = yb
            dxb, dWz, dWr, dWh = orc64.gru_seq_backward(Wz, Wr, Wh, x[b, :Lb], yb, gates, dy[b, :Lb, d * H:(d + 1) * H], reverse=rv)
            dx_ref[b, :Lb] += dxb
            dW_ref[3 * d] += dWz; dW_ref[3 * d + 1] += dWr; dW_ref[3 * d + 2] += dWh
    Wd, xd, ld = dev(W, torch.float32), dev(x, torch.float32), dev(lengths)
    y, save = s2s.gru_seq_forward(gctx, Wd, xd, lengths=ld, ndir=ndir, reverse=reverse)
    assert rel_err(y.cpu().numpy(), y_ref) < TOL
    dx, dW = s2s.gru_seq_backward(gctx, Wd, xd, y, save, dev(dy, torch.float32), lengths=ld, ndir=ndir, reverse=reverse)
    assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    assert rel_err(dW.cpu().numpy(), dW_ref) < TOL


def test_tconv_zb_known_answer_cell(s2s, gctx):
    # Attention.ipynb:123,145-154 on the GPU path: weights 1..20 row-major, input ones -> 15 40 65 90
    W = dev(np.arange(1, 21, dtype=np.float32).reshape(4, 5))
    y = s2s.tconv_zb_forward(gctx, dev(np.ones((10, 5), np.float32)), W)
    assert np.array_equal(y.cpu().numpy(), np.tile([15, 40, 65, 90], (10, 1)).astype(np.float32))
    # gradients: dx = dy W ; dW += dy^T x
    dy = dev(np.ones((10, 4), np.float32)); dW = torch.zeros(4, 5, device="cuda")
    dx = s2s.tconv_zb_backward(gctx, dev(np.ones((10, 5), np.float32)), W, dy, dW=dW)
    assert np.allclose(dx.cpu().numpy(), np.tile(np.arange(1, 21).reshape(4, 5).sum(0), (10, 1)))
    assert np.allclose(dW.cpu().numpy(), 10.0)


@pytest.mark.parametrize("H,Din,B,L,peep,reverse", [(128, 20, 4, 23, False, False), (128, 20, 4, 23, True, False), (64, 36, 5, 17, True, True),
                                                    (256, 123, 3, 12, False, True),
                                                    # persistent cluster recurrence: groups of 2, 3 and 8 (two passes) utterances per cluster
                                                    (128, 36, 20, 31, False, True), (128, 36, 40, 19, False, False), (128, 20, 70, 11, False, False),
                                                    (256, 64, 20, 15, False, False), (256, 64, 33, 9, False, True)])
def test_lstm_seq_matches_oracle(s2s, gctx, orc64, H, Din, B, L, peep, reverse):
    # nn.RNN(nn.LSTM(in, out, peepholes), reverse): LSTM.lua:6-136 (two biases per gate, full-matrix peepholes)
    rng = np.random.default_rng(H + Din + B + int(peep))
    n = orc64.lstm_param_count(Din, H, peep)
    assert s2s.lstm_param_count(Din, H, peep) == n
    P = rng.uniform(-1, 1, n) / np.sqrt(H) * 1.2
    x = rng.standard_normal((B, L, Din))
    lengths = rng.integers(max(1, L // 2), L + 1, B).astype(np.int32); lengths[0] = L
    dy = rng.standard_normal((B, L, H))
    for b in range(B):
        x[b, lengths[b]:] = 0; dy[b, lengths[b]:] = 0
    y_ref = np.zeros((B, L, H)); dx_ref = np.zeros_like(x); dP_ref = np.zeros_like(P)
    for b in range(B):
        Lb = lengths[b]
        yb, cb, ab = orc64.lstm_seq_forward(P, Din, H, peep, x[b, :Lb], reverse=reverse)
        y_ref[b, :Lb] = yb
        dxb, dPb = orc64.lstm_seq_backward(P, Din, H, peep, x[b, :Lb], yb, cb, ab, dy[b, :Lb], reverse=reverse)
        dx_ref[b, :Lb] = dxb; dP_ref += dPb
    Pd, xd, ld = dev(P, torch.float32), dev(x, torch.float32), dev(lengths)
    y, save = s2s.lstm_seq_forward(gctx, Pd, xd, H, peepholes=peep, lengths=ld, reverse=reverse)
    assert rel_err(y.cpu().numpy(), y_ref) < TOL
    dx, dP = s2s.lstm_seq_backward(gctx, Pd, xd, y, save, dev(dy, torch.float32), H, peepholes=peep, lengths=ld, reverse=reverse)
    assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    assert rel_err(dP.cpu().numpy(), dP_ref) < TOL


@pytest.mark.parametrize("B,L,S,A,KF", [(3, 70, 512, 512, 10), (2, 300, 512, 512, 10), (4, 41, 256, 512, 5)])
def test_attn_step_location_aware(s2s, gctx, B, L, S, A, KF):
    # Z[l] = q + Vh[l] + sum_j UW[j] alpha_prev[l + j - pad_left]   (Attention.lua:75-99, both convolutions folded)
    rng = np.random.default_rng(B + L + KF)
    Vh = rng.standard_normal((B, L, S)); h = rng.standard_normal((B, L, A)); q = rng.standard_normal((B, S))
    w = rng.standard_normal(S) / np.sqrt(S); uw = rng.standard_normal((KF, S))
    lengths = rng.integers(max(KF, L // 3), L + 1, B).astype(np.int32); lengths[0] = L
    ap = rng.uniform(0, 1, (B, L))
    for b in range(B):
        ap[b, lengths[b]:] = 0; ap[b] /= ap[b].sum()
    padl = (KF - 1) // 2 if KF % 2 else KF // 2
    dc = rng.standard_normal((B, A)); dain = rng.standard_normal((B, L)) * 0.1
    alpha_ref = np.zeros((B, L)); c_ref = np.zeros((B, A)); dq_ref = np.zeros((B, S)); de_ref = np.zeros((B, L)); dap_ref = np.zeros((B, L))
    for b in range(B):
        Lb = lengths[b]
        app = np.zeros(Lb + KF); app[padl:padl + Lb] = ap[b, :Lb]          # zero padding outside [0, L_b)
        loc = sum(np.outer(app[j:j + Lb], uw[j]) for j in range(KF))
        th = np.tanh(q[b][None, :] + Vh[b, :Lb] + loc)
        e = th @ w
        p = np.exp(e - e.max()); p /= p.sum()
        alpha_ref[b, :Lb] = p; c_ref[b] = p @ h[b, :Lb]
        da = h[b, :Lb] @ dc[b] + dain[b, :Lb]
        de = p * (da - p @ da)
        dZ = de[:, None] * w[None, :] * (1 - th * th)
        dq_ref[b] = dZ.sum(0); de_ref[b, :Lb] = de
        g = dZ @ uw.T                                                       # [Lb, KF]: d app[l + j] += g[l, j]
        dapp = np.zeros(Lb + KF)
        for j in range(KF):
            dapp[j:j + Lb] += g[:, j]
        dap_ref[b, :Lb] = dapp[padl:padl + Lb]
    f32 = lambda x: dev(x, torch.float32)
    dl = dev(lengths)
    alpha, c = s2s.attn_step_forward_loc(gctx, f32(Vh), f32(h), f32(q), f32(w), f32(uw), f32(ap), lengths=dl)
    assert rel_err(alpha.cpu().numpy(), alpha_ref) < TOL and rel_err(c.cpu().numpy(), c_ref) < TOL
    dq_g, de_g, dap_g = s2s.attn_step_backward_loc(gctx, f32(Vh), f32(h), f32(q), f32(w), f32(uw), f32(ap), f32(alpha_ref), f32(dc),
                                                   dalpha_in=f32(dain), lengths=dl, dalpha_prev=torch.zeros(B, L, device="cuda"))
    assert rel_err(de_g.cpu().numpy(), de_ref) < TOL and rel_err(dq_g.cpu().numpy(), dq_ref) < TOL
    assert rel_err(dap_g.cpu().numpy(), dap_ref) < TOL
