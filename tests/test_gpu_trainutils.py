"""TrainUtils (TrainUtils.lua:5-213) and the caller-side pieces of the training scripts on the GPU, through the host mirror of
the Lua shims (seq2seq-attention-asr_b200/nn.py): per-row norm constraint over a module graph, orthogonal initialisation,
the label-mask kernels, the dropout / two-stage-MLP variants of the decoder read off the caller's sub-graphs."""
import numpy as np
import pytest
import torch

from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu


def _chorowski_decoder(nn, ctx, st=64, A=256, M=8, MW=3, V=11, S=128, K=0, KF=4, dropout=None, stages=1):
    rec = nn.Sequential(ctx, nn.GRU(ctx, st, st))
    mods = [nn.Maxout(ctx, st + A, M, MW)]
    if dropout is not None:
        mods.insert(0, nn.Dropout(ctx, dropout))          # model_chorowski_baseline_dropout.lua:56
    mods.append(nn.Linear(ctx, M, M if stages == 2 else V))
    if stages == 2:                                        # librispeech/model_vgg.lua:76-80
        mods += [nn.Maxout(ctx, M, M, MW), nn.Linear(ctx, M, V)]
    mods.append(nn.LogSoftMax(ctx))
    return nn.Attention(ctx, rec, nn.Sequential(ctx, *mods), S, KF, K, st, A, V, True, 0.0)


def test_column_norm_constraint_graph_is_per_row_of_every_leaf(s2s, gctx, orc64):
    """TrainUtils.columnNormConstraintGraph (timit/timit.lua:346-348) reaches every weight matrix through apply2graph and clips
    each ROW (norm(2,2), TrainUtils.lua:63): a GRU's three gate matrices separately, the decoder's leaves one by one."""
    nn = s2s.nn
    gru = nn.GRU(gctx, 24, 32)
    gru.weight.mul_(6.0)                                   # rows well above norm 1
    W0 = gru.weight.cpu().numpy().astype(np.float64)
    rnn = nn.RNN(gctx, gru, False)
    nn.TrainUtils.columnNormConstraintGraph(rnn)
    for g in range(3):
        ref = W0[g].copy()
        assert orc64.rownorm_constraint(ref, 1.0) == 0
        assert rel_err(gru.weight[g].cpu().numpy(), ref) < 1e-6
    norms = gru.weight.view(-1, gru.weight.shape[-1]).norm(dim=1)
    assert float(norms.max()) <= 1.0 + 1e-5 and float(norms.min()) > 0.99         # clipped per row, not divided by a per-gate norm
    att = _chorowski_decoder(nn, gctx, K=3)
    att.flat.mul_(4.0)
    P0 = att.flat.cpu().numpy().astype(np.float64)
    nn.TrainUtils.columnNormConstraintGraph(att)
    ref = P0.copy()
    for (off, r, c) in s2s.param_segments(att.cfg):
        if c > 1:
            Wm = ref[off:off + r * c].reshape(r, c).copy()
            orc64.rownorm_constraint(Wm, 1.0)
            ref[off:off + r * c] = Wm.ravel()
    # biases are left alone (TrainUtils.lua:96-103 is commented out); `we` is a [1, S] row vector
    segs = dict(zip(s2s.segment_names(att.cfg), s2s.param_segments(att.cfg)))
    off, r, c = segs["we"]
    v = ref[off:off + r * c]; nrm = np.linalg.norm(v)
    if nrm >= 1.0:
        ref[off:off + r * c] = v / (nrm + 1e-8)
    assert rel_err(att.flat.cpu().numpy(), ref) < 1e-6
    with pytest.raises(s2s.S2SError):
        nn.TrainUtils.columnNormConstraint(gru)            # a 3-D holder is not a leaf: the graph walker must be used
    # LSTM leaves: every Linear of the reference's LSTM graph (LSTM.lua:25-36) is its own leaf with a bias
    lstm = nn.LSTM(gctx, 10, 16, peepholes=True)
    assert len(lstm.modules) == 11 and all(m.weight.dim() == 2 and m.bias.numel() == 16 for m in lstm.modules)
    lstm.weight.mul_(9.0)
    nn.TrainUtils.columnNormConstraintGraph(nn.RNN(gctx, lstm))
    assert max(float(m.weight.norm(dim=1).max()) for m in lstm.modules) <= 1.0 + 1e-5


@pytest.mark.parametrize("rows,cols,bias", [(64, 24, False), (24, 64, False), (448, 320, True), (62, 64, True), (256, 768, False), (1, 128, True)])
def test_orthogonalize_matches_lapack_qr(s2s, gctx, rows, cols, bias):
    """TrainUtils.orthogonalize (TrainUtils.lua:5-26): Q of the tall-orientation QR with LAPACK's sign convention (torch.qr);
    numpy.linalg.qr calls the same geqrf / orgqr."""
    rng = np.random.default_rng(rows * 1000 + cols)
    W = rng.standard_normal((rows, cols)); b = rng.standard_normal(rows) if bias else None
    w = np.concatenate([W, b[:, None]], axis=1) if bias else W
    q = np.linalg.qr(w.T)[0].T if w.shape[0] < w.shape[1] else np.linalg.qr(w)[0]
    Wd = dev(W, torch.float32); bd = dev(b, torch.float32) if bias else None
    m = s2s.nn.Param(gctx, "t", Wd, torch.zeros_like(Wd), bd, None)
    s2s.nn.TrainUtils.orthogonalize(m)
    got = np.concatenate([Wd.cpu().numpy(), bd.cpu().numpy()[:, None]], axis=1) if bias else Wd.cpu().numpy()
    assert np.abs(got - q).max() < 2e-5
    g = got.astype(np.float64)
    eye = g @ g.T if g.shape[0] < g.shape[1] else g.T @ g
    assert np.abs(eye - np.eye(eye.shape[0])).max() < 1e-5


def test_orthogonalize_graph_and_optim_config_resets(s2s, gctx):
    nn = s2s.nn
    att = _chorowski_decoder(nn, gctx, K=2)
    nn.TrainUtils.orthogonalizeGraph(att)                   # librispeech/exp0_scriptchecker.lua:49-52
    for m in att.modules:
        w = m.weight if m.weight.dim() == 2 else m.weight.view(1, -1)
        w = torch.cat([w, m.bias.view(-1, 1)], 1) if (m.bias is not None and m.bias.numel() == w.shape[0]) else w
        w = w.double()
        eye = w @ w.T if w.shape[0] < w.shape[1] else w.T @ w
        assert float((eye - torch.eye(eye.shape[0], device=eye.device, dtype=eye.dtype)).abs().max()) < 1e-4, m.name
    cfg = dict(eps=1e-8, rho=0.95)
    resets = {3: dict(eps=1e-10), 5: dict(rho=0.9, eps=1e-12)}     # optimConfigResets, timit/timit.lua:496-502
    seen = []
    for epoch in range(1, 7):
        nn.TrainUtils.optimConfigResets(cfg, resets, epoch)
        seen.append((cfg["eps"], cfg["rho"]))
    assert seen == [(1e-8, 0.95)] * 2 + [(1e-10, 0.95)] * 2 + [(1e-12, 0.9)] * 2


def test_label_mask_kernels(s2s, gctx):
    rng = np.random.default_rng(5)
    lab = rng.integers(0, 62, (7, 50)).astype(np.int32)
    oh = s2s.onehot(gctx, dev(lab), 62)                     # labelmask of timit/timit.lua:262, generated on the device
    assert np.array_equal(oh.cpu().numpy(), np.eye(62, dtype=np.float32)[lab])
    back = s2s.labels_from_onehot(gctx, oh)
    assert np.array_equal(back.cpu().numpy(), lab)
    oh[2, 3].zero_()                                        # an all-zero row (prev_y at t = 1, RNNAttention.lua:172-176) has no label
    assert int(s2s.labels_from_onehot(gctx, oh)[2, 3]) == -1


def test_decoder_variants_are_read_off_the_callers_subgraphs(s2s, gctx, orc64):
    """the `_dropout` model's nn.Dropout (model_chorowski_baseline_dropout.lua:56) and model_vgg.lua's second Maxout stage must not
    vanish: nn.Attention inspects decoder_mlp"""
    nn = s2s.nn
    rng = np.random.default_rng(9)
    att = _chorowski_decoder(nn, gctx, dropout=0.5, stages=2)
    assert att.cfg["MLP"] == 2 and att.dropout.p == 0.5
    L, T, V = 29, 4, 11
    h = rng.standard_normal((L, 256)) * 0.5
    lab = rng.integers(0, V, T); y = np.eye(V)[lab]
    P = att.flat.cpu().numpy().astype(np.float64)
    att.evaluate()                                          # nn.Dropout is the identity in evaluate mode
    out = att.forward([dev(h, torch.float32), dev(y, torch.float32)])
    ref = orc64.attention_forward(att.cfg, P, h, lab)
    assert rel_err(out.cpu().numpy(), ref["logp"]) < 1e-4
    att.training()                                          # training mode: a fresh mask per call, scaled by 1/(1-p)
    o1 = att.forward([dev(h, torch.float32), dev(y, torch.float32)]).clone()
    mask = att._args[4]
    assert mask is not None and set(np.unique(mask.cpu().numpy()).tolist()) == {0.0, 2.0}
    ref_d = orc64.attention_forward(att.cfg, P, h, lab, dropmask=mask[0].cpu().numpy().astype(np.float64))
    assert rel_err(o1.cpu().numpy(), ref_d["logp"]) < 1e-4
    o2 = att.forward([dev(h, torch.float32), dev(y, torch.float32)])
    assert not torch.equal(o1, o2)
    # an LSTM decoder is rejected loudly, not silently replaced
    with pytest.raises(s2s.S2SError):
        nn.Attention(gctx, nn.Sequential(gctx, nn.LSTM(gctx, 64, 64)), nn.Sequential(gctx, nn.Maxout(gctx, 320, 8, 3), nn.Linear(gctx, 8, 11)),
                     128, 4, 0, 64, 256, 11, False, 0.0)
    # AdaptiveWeightNoise: sigma_init defaults to 1 as in the reference (AdaptiveWeightNoise.lua:13) -> s = log 1 = 0
    awn = nn.AdaptiveWeightNoise(gctx, torch.zeros(16, device="cuda"))
    assert float(awn.weight[16:].abs().max()) == 0.0
    # reset(stdv) draws from U(+-stdv sqrt 3) (LinearZeroBias.lua:13-14)
    lin = nn.LinearZeroBias(gctx, 64, 64); lin.reset(0.1)
    assert 0.15 < float(lin.weight.abs().max()) <= 0.1 * 3 ** 0.5 + 1e-6
