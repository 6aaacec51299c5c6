"""VGG front-end (librispeech/model_vgg.lua:23-54) on the GPU against the numpy restatement (oracle/vgg.py)."""
import numpy as np
import pytest
import torch

from oracle import vgg
from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("cfg,B,T,F", [(dict(C1=8, C2=12, HID=40, OUT=16), 3, 22, 24), (dict(C1=16, C2=32, HID=64, OUT=32), 2, 37, 40),
                                       (dict(C1=64, C2=128, HID=256, OUT=128), 5, 48, 40)])
def test_vgg_forward_backward_matches_oracle(s2s, gctx, cfg, B, T, F):
    rng = np.random.default_rng(B + T + F)
    P = vgg.init_params(cfg, F, seed=T) * 1.5
    assert s2s.vgg_param_count(cfg, F) == P.size == vgg.param_count(cfg, F)
    X = rng.standard_normal((B, 3, T, F))
    L = vgg.out_len(T)
    dh = rng.standard_normal((B, L, cfg["OUT"]))
    h_ref = np.zeros((B, L, cfg["OUT"])); dP_ref = np.zeros_like(P); dX_ref = np.zeros_like(X)
    for b in range(B):
        hb, cache = vgg.forward(cfg, P, X[b])
        h_ref[b] = hb
        dPb, dXb = vgg.backward(cfg, P, cache, dh[b])
        dP_ref += dPb; dX_ref[b] = dXb
    Pd, Xd = dev(P, torch.float32), dev(X, torch.float32)
    h = s2s.vgg_forward(gctx, cfg, Pd, Xd)
    assert rel_err(h.cpu().numpy(), h_ref) < TOL
    dP, dX = s2s.vgg_backward(gctx, cfg, Pd, Xd, dev(dh, torch.float32), need_dx=True)
    segs, o = vgg.segments(cfg, F), 0
    dPg = dP.cpu().numpy()
    for name, shape in segs:
        n = int(np.prod(shape))
        assert rel_err(dPg[o:o + n], dP_ref[o:o + n]) < TOL, name
        o += n
    assert rel_err(dX.cpu().numpy(), dX_ref) < TOL


def test_vgg_padded_batch_equals_per_utterance(s2s, gctx):
    # the reference runs one utterance at a time; a padded batch must give every utterance the result it gets alone:
    # annotations l < L_b = (T_b - 8) // 2 only see input frames < T_b, and zero dh beyond L_b keeps the gradients exact
    cfg = dict(C1=32, C2=64, HID=48, OUT=32)          # multiples of 32: the implicit-GEMM path
    F, T, Tb = 24, 44, (44, 31, 26)
    B = len(Tb)
    rng = np.random.default_rng(12)
    P = vgg.init_params(cfg, F, seed=2) * 1.5
    X = rng.standard_normal((B, 3, T, F))
    L = vgg.out_len(T)
    dh = rng.standard_normal((B, L, cfg["OUT"]))
    h_ref = np.zeros((B, L, cfg["OUT"])); dP_ref = np.zeros_like(P); dX_ref = np.zeros_like(X)
    for b in range(B):
        X[b, :, Tb[b]:] = 0
        Lb = vgg.out_len(Tb[b])
        dh[b, Lb:] = 0
        hb, cache = vgg.forward(cfg, P, X[b, :, :Tb[b]])
        h_ref[b, :Lb] = hb
        dPb, dXb = vgg.backward(cfg, P, cache, dh[b, :Lb])
        dP_ref += dPb; dX_ref[b, :, :Tb[b]] = dXb
    Pd, Xd = dev(P, torch.float32), dev(X, torch.float32)
    h = s2s.vgg_forward(gctx, cfg, Pd, Xd).cpu().numpy()
    for b in range(B):
        Lb = vgg.out_len(Tb[b])
        assert rel_err(h[b, :Lb], h_ref[b, :Lb]) < TOL
    dP, dX = s2s.vgg_backward(gctx, cfg, Pd, Xd, dev(dh, torch.float32), need_dx=True)
    assert rel_err(dP.cpu().numpy(), dP_ref) < TOL
    dXg = dX.cpu().numpy()
    assert rel_err(dXg, dX_ref) < TOL
    for b in range(B):
        assert np.abs(dXg[b, :, Tb[b]:]).max() == 0.0 if Tb[b] < T else True
