"""tcgen05 (3xTF32) GEMM against float64: must hold the fp32 parity bound with margin."""
import numpy as np
import pytest
import torch

from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 128, 64), (300, 200, 96), (1000, 1536, 512), (9600, 512, 512), (130, 70, 36)])
def test_gemm_tcgen05_3xtf32(s2s, gctx, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = 0.5 * (A.astype(np.float64) @ B.astype(np.float64).T) + 2.0 * C0 + bias
    Cd = dev(C0)
    s2s.gemm(gctx, dev(A), dev(B), tA=False, tB=True, alpha=0.5, beta=2.0, C_out=Cd, bias=dev(bias), impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


def test_gemm_tcgen05_strided_operands(s2s, gctx):
    # the encoder projection reads the x-columns of the GRU weights: B = W[:, H:] with row pitch H + Din
    rng = np.random.default_rng(5)
    M, N, K, H = 640, 384, 256, 128
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = rng.standard_normal((N, H + K)).astype(np.float32)
    ref = A.astype(np.float64) @ W[:, H:].astype(np.float64).T
    Wd = dev(W)
    Cd = torch.zeros(M, N, device="cuda")
    s2s.gemm(gctx, dev(A), Wd[:, H:], tA=False, tB=True, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


@pytest.mark.parametrize("beta", [0.0, 1.0])
@pytest.mark.parametrize("M,N,K", [(1000, 1536, 512), (9600, 512, 512), (640, 300, 1000)])
def test_gemm_tcgen05_streamk_tail(s2s, gctx, M, N, K, beta):
    # tile counts that do not fill a round of 148 CTAs: the left-over tiles are cut along K and added with atomics
    rng = np.random.default_rng(M + N + K + int(beta))
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T + beta * C0 + bias
    Cd = dev(C0)
    s2s.gemm(gctx, dev(A), dev(B), tA=False, tB=True, alpha=1.0, beta=beta, C_out=Cd, bias=dev(bias), impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


@pytest.mark.parametrize("M,N,K", [(9600, 512, 768), (1600, 123, 1536)])
def test_gemm_tcgen05_nn_data_gradient(s2s, gctx, M, N, K):
    # dX = dA . W : B is [K, N] (N-contiguous) and is transposed + split by the pre-pass
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    Cd = torch.full((M, N), 7.0, device="cuda")
    s2s.gemm(gctx, dev(A), dev(B), tA=False, tB=False, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


@pytest.mark.parametrize("M,N,K", [(1536, 512, 9600), (768, 123, 9600), (256, 256, 1600)])
def test_gemm_tcgen05_tn_weight_gradient(s2s, gctx, M, N, K):
    # dW += dA^T X : both operands are [K, .]; K = B*L is long, the output small -> all tiles are stream-K shares
    rng = np.random.default_rng(M + K)
    A = rng.standard_normal((K, M)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = A.astype(np.float64).T @ B.astype(np.float64) + C0
    Cd = dev(C0)
    s2s.gemm(gctx, dev(A), dev(B), tA=True, tB=False, alpha=1.0, beta=1.0, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


def test_gemm_tcgen05_unaligned_k(s2s, gctx):
    # layer-1 encoder projection: K = 123 with pitch 123 (not a multiple of 4) -> both parts are materialised
    rng = np.random.default_rng(9)
    M, N, K = 2000, 768, 123
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = rng.standard_normal((N, 256 + K)).astype(np.float32)
    ref = A.astype(np.float64) @ W[:, 256:].astype(np.float64).T
    Cd = torch.zeros(M, N, device="cuda")
    s2s.gemm(gctx, dev(A), dev(W)[:, 256:], tA=False, tB=True, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5


@pytest.mark.parametrize("form,M,N,K", [("NT", 5000, 64, 576), ("NT", 3000, 128, 1152), ("NT", 2000, 100, 96), ("NN", 4000, 27, 64),
                                        ("TN", 64, 576, 20000), ("TN", 128, 123, 9600)])
def test_gemm_tcgen05_narrow_tiles(s2s, gctx, form, M, N, K):
    # N <= 64 / <= 128 run on 64- / 128-wide tiles (the 64- and 128-plane convolutions of the VGG front-end)
    rng = np.random.default_rng(M + N + K)
    tA, tB = form[0] == "T", form[1] == "T"
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    beta = 1.0 if tA else 0.0
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64) + beta * C0
    Cd = dev(C0)
    s2s.gemm(gctx, dev(A), dev(B), tA=tA, tB=tB, beta=beta, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    assert rel_err(Cd.cpu().numpy(), ref) < 2e-5
