"""CPU restatement (numpy, float64) of the VGG front-end of librispeech/model_vgg.lua:23-54.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and benchmarks run by hand); the product package never imports it.
Parity status: **unpinned** -- the reference holds no golden vector for this encoder and its arithmetic lives in
un-vendored Torch7 `nn` modules, restated here from their published semantics:
  * nn.SpatialConvolutionMM(nIn, nOut, kW=3, kH=3): valid cross-correlation, weight [nOut, nIn*kH*kW] with the
    (plane, kh, kw) column order of the unfolded input, bias [nOut]                       (model_vgg.lua:24-33)
  * nn.ReLU; nn.SpatialMaxPooling(kW, kH, dW, dH) in floor mode, first maximum wins       (model_vgg.lua:28,33)
  * nn.Transpose2({1,2},3): nFeat x L x H -> L x nFeat x H; nn.View(-1, nFeat*H)          (model_vgg.lua:45-46)
  * nn.TemporalConvolution(in, out, 1) == Linear with bias applied to every frame         (model_vgg.lua:47-54)
tests/test_oracle_vgg.py checks this file against torch.nn.functional.conv2d / max_pool2d + autograd in float64.
Input per utterance: X [3, T, F] (planes, time, frequency); output [L, OUT] with L = floor((T - 8) / 2).
"""
import numpy as np

# (conv1/2 planes, conv3/4 planes, hidden width of the 1x1 stack, annotation depth): model_vgg.lua:24-54 defaults
VGG_LIBRISPEECH = dict(C1=64, C2=128, HID=2048, OUT=512)


def out_freq(F):
    H = F - 4
    H = H // 2
    H = H - 4
    return H // 2                                                       # model_vgg.lua:38-43


def out_len(T):
    return (T - 8) // 2                                                 # model_vgg.lua:35-36


def segments(cfg, F):
    """(name, shape) of every parameter in the order encoder:parameters() yields them."""
    C1, C2, HID, OUT = cfg["C1"], cfg["C2"], cfg["HID"], cfg["OUT"]
    view = C2 * out_freq(F)
    return [("conv1.W", (C1, 3 * 9)), ("conv1.b", (C1,)), ("conv2.W", (C1, C1 * 9)), ("conv2.b", (C1,)),
            ("conv3.W", (C2, C1 * 9)), ("conv3.b", (C2,)), ("conv4.W", (C2, C2 * 9)), ("conv4.b", (C2,)),
            ("t1.W", (HID, view)), ("t1.b", (HID,)), ("t2.W", (HID, HID)), ("t2.b", (HID,)),
            ("t3.W", (HID, HID)), ("t3.b", (HID,)), ("t4.W", (OUT, HID)), ("t4.b", (OUT,))]


def param_count(cfg, F):
    return int(sum(int(np.prod(s)) for _, s in segments(cfg, F)))


def unflatten(cfg, F, P):
    out, o = {}, 0
    for name, shape in segments(cfg, F):
        n = int(np.prod(shape))
        out[name] = P[o:o + n].reshape(shape)
        o += n
    assert o == P.size
    return out


def init_params(cfg, F, seed=0):
    """uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) like the modules' reset()"""
    rng = np.random.default_rng(seed)
    parts = []
    segs = segments(cfg, F)
    for i in range(0, len(segs), 2):
        fan_in = segs[i][1][1]
        for _, shape in segs[i:i + 2]:
            parts.append(rng.uniform(-1, 1, int(np.prod(shape))) / np.sqrt(fan_in))
    return np.concatenate(parts)


def _unfold(x):
    """x [C, Hh, Ww] -> [Hh-2, Ww-2, C*9] with column order (plane, kh, kw)"""
    C, Hh, Ww = x.shape
    w = np.lib.stride_tricks.sliding_window_view(x, (3, 3), axis=(1, 2))       # [C, Hh-2, Ww-2, 3, 3]
    return np.ascontiguousarray(w.transpose(1, 2, 0, 3, 4)).reshape(Hh - 2, Ww - 2, C * 9)


def conv_fwd(x, W, b):
    col = _unfold(x)
    y = col @ W.T + b                                                           # [Ho, Wo, Cout]
    return np.ascontiguousarray(y.transpose(2, 0, 1)), col


def conv_bwd(x_shape, col, W, dy):
    """dy [Cout, Ho, Wo] -> dx [C, Hh, Ww], dW, db"""
    C, Hh, Ww = x_shape
    d = dy.transpose(1, 2, 0)                                                   # [Ho, Wo, Cout]
    dW = d.reshape(-1, d.shape[2]).T @ col.reshape(-1, col.shape[2])
    db = d.sum((0, 1))
    dcol = (d @ W).reshape(Hh - 2, Ww - 2, C, 3, 3)
    dx = np.zeros(x_shape)
    for kh in range(3):
        for kw in range(3):
            dx[:, kh:kh + Hh - 2, kw:kw + Ww - 2] += dcol[:, :, :, kh, kw].transpose(2, 0, 1)
    return dx, dW, db


def pool_fwd(x, kW, kH):
    """SpatialMaxPooling(kW, kH, kW, kH), floor mode: x [C, Hh, Ww] -> y, argmax (first maximum in (kh, kw) scan order)"""
    C, Hh, Ww = x.shape
    Ho, Wo = Hh // kH, Ww // kW
    v = x[:, :Ho * kH, :Wo * kW].reshape(C, Ho, kH, Wo, kW).transpose(0, 1, 3, 2, 4).reshape(C, Ho, Wo, kH * kW)
    idx = v.argmax(3)                                                           # numpy argmax returns the first maximum
    return np.take_along_axis(v, idx[..., None], 3)[..., 0], idx


def pool_bwd(x_shape, idx, dy, kW, kH):
    C, Hh, Ww = x_shape
    Ho, Wo = dy.shape[1:]
    g = np.zeros((C, Ho, Wo, kH * kW))
    np.put_along_axis(g, idx[..., None], dy[..., None], 3)
    dx = np.zeros(x_shape)
    dx[:, :Ho * kH, :Wo * kW] = g.reshape(C, Ho, Wo, kH, kW).transpose(0, 1, 3, 2, 4).reshape(C, Ho * kH, Wo * kW)
    return dx


def _pool_margin(x, kW, kH):
    """smallest (max - runner-up) over the pooling windows whose maximum is positive, relative to the tensor's RMS"""
    C, Hh, Ww = x.shape
    Ho, Wo = Hh // kH, Ww // kW
    v = np.sort(x[:, :Ho * kH, :Wo * kW].reshape(C, Ho, kH, Wo, kW).transpose(0, 1, 3, 2, 4).reshape(C, Ho, Wo, kH * kW), -1)
    act = v[..., -1] > 0
    return float((v[..., -1] - v[..., -2])[act].min() / np.sqrt(np.mean(x * x) + 1e-300)) if act.any() else np.inf


def forward(cfg, P, X, margins=None):
    """X [3, T, F] -> (h [L, OUT], cache).  margins (a list) collects, per ReLU and per pooling, how close the closest
    decision of this utterance is to flipping (|pre-activation| / RMS, window max - runner-up / RMS): the network is
    piecewise linear, and a float32 evaluation that lands on the other side of such a decision has different gradients."""
    p = unflatten(cfg, X.shape[2], P)
    c = {"x0": X}

    def relu(z):
        if margins is not None:
            margins.append(float(np.abs(z).min() / np.sqrt(np.mean(z * z) + 1e-300)))
        return np.maximum(z, 0)

    a, c["col1"] = conv_fwd(X, p["conv1.W"], p["conv1.b"]); a = relu(a); c["a1"] = a
    a, c["col2"] = conv_fwd(a, p["conv2.W"], p["conv2.b"]); a = relu(a); c["a2"] = a
    if margins is not None:
        margins.append(_pool_margin(a, 2, 1))
    a, c["i1"] = pool_fwd(a, 2, 1); c["p1"] = a                                  # SpatialMaxPooling(2, 1, 2, 1)
    a, c["col3"] = conv_fwd(a, p["conv3.W"], p["conv3.b"]); a = relu(a); c["a3"] = a
    a, c["col4"] = conv_fwd(a, p["conv4.W"], p["conv4.b"]); a = relu(a); c["a4"] = a
    if margins is not None:
        margins.append(_pool_margin(a, 2, 2))
    a, c["i2"] = pool_fwd(a, 2, 2); c["p2"] = a                                  # SpatialMaxPooling(2, 2, 2, 2)
    f = np.ascontiguousarray(a.transpose(1, 0, 2)).reshape(a.shape[1], -1)        # Transpose2 + View: [L, nFeat*H]
    c["f0"] = f
    for k in (1, 2, 3, 4):
        f = relu(f @ p[f"t{k}.W"].T + p[f"t{k}.b"])
        c[f"f{k}"] = f
    return f, c


def backward(cfg, P, c, dh):
    """dh [L, OUT] -> (dP flat, dX [3, T, F])"""
    F = c["x0"].shape[2]
    p = unflatten(cfg, F, P)
    g = {}
    d = dh
    for k in (4, 3, 2, 1):
        d = d * (c[f"f{k}"] > 0)
        g[f"t{k}.W"] = d.T @ c[f"f{k - 1}"]; g[f"t{k}.b"] = d.sum(0)
        d = d @ p[f"t{k}.W"]
    C2, L, Hq = c["p2"].shape
    d = d.reshape(L, C2, Hq).transpose(1, 0, 2)
    d = pool_bwd(c["a4"].shape, c["i2"], d, 2, 2)
    d = d * (c["a4"] > 0); d, g["conv4.W"], g["conv4.b"] = conv_bwd(c["a3"].shape, c["col4"], p["conv4.W"], d)
    d = d * (c["a3"] > 0); d, g["conv3.W"], g["conv3.b"] = conv_bwd(c["p1"].shape, c["col3"], p["conv3.W"], d)
    d = pool_bwd(c["a2"].shape, c["i1"], d, 2, 1)
    d = d * (c["a2"] > 0); d, g["conv2.W"], g["conv2.b"] = conv_bwd(c["a1"].shape, c["col2"], p["conv2.W"], d)
    d = d * (c["a1"] > 0); d, g["conv1.W"], g["conv1.b"] = conv_bwd(c["x0"].shape, c["col1"], p["conv1.W"], d)
    dP = np.concatenate([g[name].reshape(-1) for name, _ in segments(cfg, F)])
    return dP, d
