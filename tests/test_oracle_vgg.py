"""The numpy restatement of the VGG front-end (oracle/vgg.py) against an independent float64 torch restatement
(conv2d / max_pool2d / linear + autograd).  Neither is the reference itself: parity for this encoder is unpinned."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from oracle import vgg

CFG = dict(C1=6, C2=10, HID=24, OUT=12)


def torch_forward(cfg, P, X):
    F = X.shape[2]
    p = {k: torch.tensor(v) for k, v in vgg.unflatten(cfg, F, P.detach().numpy()).items()}
    # rebuild differentiable views of P
    out, o = {}, 0
    for name, shape in vgg.segments(cfg, F):
        n = int(np.prod(shape)); out[name] = P[o:o + n].view(shape); o += n
    p = out
    def conv(a, W, b):
        Cout = W.shape[0]
        return Fn.conv2d(a[None], W.view(Cout, -1, 3, 3), b)[0]
    a = torch.relu(conv(X, p["conv1.W"], p["conv1.b"]))
    a = torch.relu(conv(a, p["conv2.W"], p["conv2.b"]))
    a = Fn.max_pool2d(a[None], kernel_size=(1, 2), stride=(1, 2))[0]
    a = torch.relu(conv(a, p["conv3.W"], p["conv3.b"]))
    a = torch.relu(conv(a, p["conv4.W"], p["conv4.b"]))
    a = Fn.max_pool2d(a[None], kernel_size=(2, 2), stride=(2, 2))[0]
    f = a.transpose(0, 1).reshape(a.shape[1], -1)
    for k in (1, 2, 3, 4):
        f = torch.relu(f @ p[f"t{k}.W"].T + p[f"t{k}.b"])
    return f


@pytest.mark.parametrize("T,F", [(20, 24), (27, 40), (13, 21)])
def test_vgg_oracle_matches_torch_autograd(T, F):
    rng = np.random.default_rng(T + F)
    P = vgg.init_params(CFG, F, seed=T) * 2.0
    X = rng.standard_normal((3, T, F))
    h, cache = vgg.forward(CFG, P, X)
    assert h.shape == (vgg.out_len(T), CFG["OUT"])
    dh = rng.standard_normal(h.shape)
    dP, dX = vgg.backward(CFG, P, cache, dh)
    Pt = torch.tensor(P, requires_grad=True); Xt = torch.tensor(X, requires_grad=True)
    ht = torch_forward(CFG, Pt, Xt)
    (ht * torch.tensor(dh)).sum().backward()
    assert np.allclose(h, ht.detach().numpy(), rtol=1e-10, atol=1e-12)
    assert np.allclose(dP, Pt.grad.numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(dX, Xt.grad.numpy(), rtol=1e-9, atol=1e-11)
    assert (h > 0).mean() > 0.05          # the test is not vacuous: some units are active


def test_decision_margins_are_reported():
    # forward(..., margins=[]) collects one margin per ReLU (8) and per pooling (2): positive, and tiny at realistic sizes --
    # the reason the GPU parity tests pick their data by margin (tests/test_gpu_vgg.py)
    cfg = dict(C1=8, C2=12, HID=40, OUT=16)
    rng = np.random.default_rng(0)
    P = vgg.init_params(cfg, 24, seed=0) * 1.5
    X = rng.standard_normal((3, 22, 24))
    m = []
    h, _ = vgg.forward(cfg, P, X, margins=m)
    h2, _ = vgg.forward(cfg, P, X)
    assert len(m) == 10 and all(v > 0 for v in m) and min(m) < 1e-2
    assert np.array_equal(h, h2)
