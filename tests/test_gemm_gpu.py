"""GEMM kernels against torch fp64 matmul, through the C-ABI test hook gmvae_debug_gemm."""
import pytest
import torch

from tests.helpers import CONFIGS, make_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = make_engine(CONFIGS["tiny_vae"], "bf16")
    yield e
    e.close()


def _ref(A, B, ta, tb):
    A64, B64 = A.double(), B.double()
    return (A64.t() if ta else A64) @ (B64.t() if tb else B64)


@pytest.mark.parametrize("M,N,K,ta,tb,split", [
    (37, 50, 19, False, False, 1), (100, 512, 784, False, False, 1), (100, 10, 512, False, True, 1),
    (784, 512, 100, True, False, 1), (10, 128, 4096, True, False, 16), (65, 67, 1000, True, True, 3),
])
def test_simt_gemm(eng, M, N, K, ta, tb, split):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    C = eng.debug_gemm(0, A, B, ta, tb, split).cpu().double()
    R = _ref(A, B, ta, tb)
    assert ((C - R).norm() / R.norm()).item() < 2e-6


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


# K-major x K-major: forward / dgrad shapes.  A [M,K], B stored [N,K].
@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (128, 64, 128), (100, 512, 784), (256, 784, 512), (300, 128, 512), (100, 64, 512),
    (129, 1024, 64), (16384, 512, 512), (100, 112, 72), (1000, 256, 1024), (16384, 784, 512), (4000, 10, 512),
    (4000, 512, 10), (3000, 24, 40), (20000, 128, 10),
])
def test_tc_gemm_kmajor(eng, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = _bf16(torch.randn(M, K, generator=g)); B = _bf16(torch.randn(N, K, generator=g))
    C = eng.debug_gemm(1, A, B, False, True).cpu().double()
    R = _ref(A, B, False, True)
    assert ((C - R).norm() / R.norm()).item() < 1e-5     # bf16 inputs are exact; only fp32 accumulation differs


# MN-major x MN-major: weight-gradient shapes.  A stored [K,M], B stored [K,N]; K = batch.
@pytest.mark.parametrize("M,N,K,split", [
    (128, 128, 64, 1), (128, 64, 128, 1), (512, 512, 100, 1), (784, 512, 100, 2), (512, 784, 256, 3),
    (64, 512, 1000, 4), (512, 128, 16384, 37), (200, 136, 333, 2), (512, 512, 16384, 18), (784, 512, 16384, 10),
    (10, 512, 5000, 8), (512, 10, 5000, 8),
])
def test_tc_gemm_mnmajor(eng, M, N, K, split):
    g = torch.Generator().manual_seed(M + N + K)
    A = _bf16(torch.randn(K, M, generator=g)); B = _bf16(torch.randn(K, N, generator=g))
    C = eng.debug_gemm(1, A, B, True, False, split).cpu().double()
    R = _ref(A, B, True, False)
    assert ((C - R).norm() / R.norm()).item() < 1e-5
