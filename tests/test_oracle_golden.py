"""Pins the oracle's primitives against the ONLY reproducible known-answer vectors the reference
holds: four deterministic iTorch notebook cells.  Everything else on the hot path
is 'parity unpinned' (no Torch7 in the image, no vectors in the reference)."""
import numpy as np
import pytest


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_temporal_convolution_cell(prec, orc32, orc64):
    # Attention.ipynb:123,145-154: TemporalConvolution(5->4, kW=1), weight = 1..20 row-major, bias 0,
    # input ones(10,5) -> every output row is 15 40 65 90.  Pins the [out, kW*in] weight layout.
    o = orc32 if prec == "f32" else orc64
    W = np.arange(1, 21).reshape(4, 5)
    y = o.tconv_forward(np.ones((10, 5)), W, np.zeros(4), kW=1)
    assert y.shape == (10, 4)
    assert np.array_equal(y, np.tile([15, 40, 65, 90], (10, 1)))


def test_temporal_convolution_kw_gt1(orc64):
    # frame-major unfolding: out[t] = W . vec(in[t:t+kW, :]) (cross-correlation, no flip)
    x = np.arange(12, dtype=np.float64).reshape(6, 2)
    W = np.arange(1, 13, dtype=np.float64).reshape(2, 6)  # kW=3, in=2
    y = orc64.tconv_forward(x, W, np.array([0.5, -0.5]), kW=3)
    ref = np.stack([W @ x[t:t + 3].reshape(-1) + np.array([0.5, -0.5]) for t in range(4)])
    assert np.allclose(y, ref)


def test_padding_cell(orc64):
    # Attention.ipynb:257: Padding(1,-2,2) then Padding(1,2,2) on ones(10,1) -> 0 0 1x10 0 0 ("negative pad = left")
    x = np.ones((10, 1))
    y = orc64.padding(orc64.padding(x, -2), 2)
    assert np.array_equal(y[:, 0], np.array([0, 0] + [1] * 10 + [0, 0], dtype=np.float64))


def test_addbias_gradbias_cell(orc64):
    # Attention.ipynb:725-752: nn.AddBias gradBias = 9 (SGD, L=9) / 27 (batch 3x9)
    assert orc64.addbias_gradbias(np.ones(9)) == 9.0
    assert orc64.addbias_gradbias(np.ones((3, 9))) == 27.0


def test_mm_cell(orc64):
    # Attention.ipynb:918-958: nn.MM of ones[1,10] x ones[10,A] -> 10 everywhere
    c = orc64.mm(np.ones((1, 10)), np.ones((10, 6)))
    assert np.array_equal(c, np.full((1, 6), 10.0))
