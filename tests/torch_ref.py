"""Independent float64 autograd restatement of the reference model (TEST ONLY).

Built from torch ops + autograd so that every gradient of the hand-written backward in
oracle/s2s_oracle.c (and of the CUDA kernels) is checked against an independent derivation.
Follows the reference graph: timit/model_chorowski_baseline.lua:20-75, Attention.lua:39-211,
GRU.lua:22-30, LSTM.lua:25-58, Maxout.lua:14-19, MonotonicAlignment.lua:19-77.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from oracle.oracle import Oracle, segment_names


def unflatten(cfg, P):
    """flat tensor -> dict name -> view, following the oracle layout"""
    o = Oracle("f64")
    segs = o.param_segments(cfg)
    out = {}
    for (off, rows, cols), name in zip(segs, segment_names(cfg)):
        v = P[off:off + rows * cols]
        out[name] = v.view(rows, cols) if (cols > 1 or name == "we") else v
    return out


def gru_step(Wz, Wr, Wh, x, hp):
    hx = torch.cat([hp, x])
    z = torch.sigmoid(Wz @ hx)
    r = torch.sigmoid(Wr @ hx)
    hc = torch.tanh(Wh @ torch.cat([r * hp, x]))
    return (1 - z) * hp + z * hc


def gru_seq(Wz, Wr, Wh, x, reverse):
    L = x.shape[0]
    H = Wz.shape[0]
    h = torch.zeros(H, dtype=x.dtype)
    out = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:
        h = gru_step(Wz, Wr, Wh, x[t], h)
        out[t] = h
    return torch.stack(out)


def lstm_unpack(P, inp, out, peep):
    gates = []
    o = 0
    for k in range(4):
        pk = peep and k != 2
        Wx = P[o:o + out * inp].view(out, inp); o += out * inp
        bx = P[o:o + out]; o += out
        Wh = P[o:o + out * out].view(out, out); o += out * out
        bh = P[o:o + out]; o += out
        Wc = bc = None
        if pk:
            Wc = P[o:o + out * out].view(out, out); o += out * out
            bc = P[o:o + out]; o += out
        gates.append((Wx, bx, Wh, bh, Wc, bc))
    assert o == P.numel()
    return gates


def lstm_step(gates, x, hp, cp):
    def pre(g, c):
        Wx, bx, Wh, bh, Wc, bc = g
        a = Wx @ x + bx + Wh @ hp + bh
        if Wc is not None:
            a = a + Wc @ c + bc
        return a
    i = torch.sigmoid(pre(gates[0], cp))
    f = torch.sigmoid(pre(gates[1], cp))
    g = torch.tanh(pre(gates[2], cp))
    cn = f * cp + i * g
    o = torch.sigmoid(pre(gates[3], cn))
    return o * torch.tanh(cn), cn


def lstm_seq(P, inp, out, peep, x, reverse):
    gates = lstm_unpack(P, inp, out, peep)
    L = x.shape[0]
    h = torch.zeros(out, dtype=x.dtype); c = torch.zeros(out, dtype=x.dtype)
    ys = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:
        h, c = lstm_step(gates, x[t], h, c)
        ys[t] = h
    return torch.stack(ys)


def encoder(cfg, p, X):
    a = X
    for l in range(cfg["NL"]):
        f = gru_seq(p[f"enc{l}f.Wz"], p[f"enc{l}f.Wr"], p[f"enc{l}f.Wh"], a, False)
        r = gru_seq(p[f"enc{l}r.Wz"], p[f"enc{l}r.Wr"], p[f"enc{l}r.Wh"], a, True)
        a = torch.cat([f, r], dim=1)
    return a


class _Mono(torch.autograd.Function):
    """MonotonicAlignment.lua: identity forward; backward injects +-lambda*(L+1-l)*1[penalty>0]"""

    @staticmethod
    def forward(ctx, alpha, prev, lam):
        pen = torch.clamp((torch.cumsum(alpha, 0) - torch.cumsum(prev, 0)).sum(), min=0)
        ctx.ind = float(pen > 0); ctx.lam = lam; ctx.L = alpha.shape[0]
        return alpha.clone()

    @staticmethod
    def backward(ctx, g):
        L = ctx.L
        gd = ctx.lam * ctx.ind * (L + 1 - torch.arange(1, L + 1, dtype=g.dtype))
        return g + gd, -gd, None


def decoder(cfg, p, h, labels, lam=0.0, dropmask=None):
    """teacher-forced nn.Attention forward; returns logp [T,V], alpha [T,L], aux"""
    S, ST, V, K, KF, M, MW = (cfg[k] for k in ("S", "ST", "V", "K", "KF", "M", "MW"))
    L = h.shape[0]; T = len(labels)
    Vh = h @ p["WV"].t()
    alpha = torch.zeros(L, dtype=h.dtype); s = torch.zeros(ST, dtype=h.dtype)
    logps, alphas, ss, cs, qs = [], [], [], [], []
    if K > 0:
        pl = (KF - 1) // 2 if KF % 2 == 1 else KF // 2
        pr = pl if KF % 2 == 1 else pl - 1
    for t in range(T):
        y = torch.zeros(V, dtype=h.dtype)
        if t > 0:
            y[labels[t - 1]] = 1
        q = p["Ws"] @ s + p["bs"]
        Z = q[None, :] + Vh
        if K > 0:
            ap = Fn.pad(alpha, (pl, pr))
            Fm = Fn.conv1d(ap[None, None, :], p["WF"][:, None, :], p["bF"])[0].t()  # [L,K]
            Z = Z + Fm @ p["U"].t()
        e = torch.tanh(Z) @ p["we"].view(-1)
        a_new = torch.softmax(e, 0)
        a_new = _Mono.apply(a_new, alpha, lam)
        c = a_new @ h
        y_in = p["Wy"] @ y + p["by"]
        c_in = p["Wc"] @ c + p["bc"]
        u = p["Wj"] @ torch.cat([c_in, y_in]) + p["bj"]
        s_new = gru_step(p["Gz"], p["Gr"], p["Gh"], u, s)
        sc = torch.cat([s_new, c])
        if dropmask is not None:
            sc = sc * dropmask[t]
        m = p["Wm"] @ sc + p["bm"]
        mo = m.view(M, MW).max(dim=1).values
        if cfg.get("MLP", 1) == 2:                 # Linear(M,M) -> Maxout(M,M,7)   (librispeech/model_vgg.lua:78-79)
            m2 = p["Wm2"] @ (p["Wl"] @ mo + p["bl"]) + p["bm2"]
            mo = m2.view(M, MW).max(dim=1).values
        logp = torch.log_softmax(p["Wo"] @ mo + p["bo"], 0)
        logps.append(logp); alphas.append(a_new); ss.append(s_new); cs.append(c); qs.append(q)
        alpha, s = a_new, s_new
    return torch.stack(logps), torch.stack(alphas), dict(s=torch.stack(ss), c=torch.stack(cs), q=torch.stack(qs), Vh=Vh)


def model_fwdbwd(cfg, P, X, lengths, labels, tlens, lam=0.0, dropmask=None, normalize_nll=False, normalize_grad=False):
    """numpy in / numpy out; float64; per-utterance loop with gradient accumulation (timit.lua:240-289)"""
    Pt = torch.tensor(np.asarray(P, dtype=np.float64), requires_grad=True)
    p = unflatten(cfg, Pt)
    B = X.shape[0]
    nll = np.zeros(B); logps = []; alphas = []; annots = []; dXs = []
    for b in range(B):
        L = int(lengths[b]) if lengths is not None else X.shape[1]
        T = int(tlens[b]) if tlens is not None else labels.shape[1]
        x = torch.tensor(np.asarray(X[b, :L], dtype=np.float64), requires_grad=True)
        lab = [int(v) for v in labels[b, :T]]
        h = encoder(cfg, p, x)
        dm = None if dropmask is None else torch.tensor(np.asarray(dropmask[b, :T], dtype=np.float64))
        logp, alpha, _ = decoder(cfg, p, h, lab, lam, dm)
        mask = torch.zeros_like(logp); mask[torch.arange(T), torch.tensor(lab)] = 1
        nll[b] = float(-(mask * logp).sum().detach()) / (T if normalize_nll else 1)
        dlogp = -mask / (T if normalize_grad else 1)
        logp.backward(dlogp)
        logps.append(logp.detach().numpy()); alphas.append(alpha.detach().numpy()); annots.append(h.detach().numpy()); dXs.append(x.grad.numpy())
    return dict(nll=nll, G=Pt.grad.numpy(), logp=logps, alpha=alphas, annot=annots, dX=dXs)
