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
