"""tcgen05 split-fp16 GEMM kernel (kraken_b200/csrc/gemm_tc.cuh) against float64 numpy and against the CUDA-core kernel."""
import ctypes as C

import numpy as np
import pytest

import kraken_b200 as kb
from kraken_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def gemm(a, b, bias, use_tc):
    M, K = a.shape
    N = b.shape[0]
    c = np.empty((M, N), np.float32)
    check(lib.kb_debug_gemm(a.ctypes.data, b.ctypes.data, bias.ctypes.data if bias is not None else None, c.ctypes.data, M, N, K, int(use_tc), 0))
    return c


@pytest.mark.parametrize('M,N,K', [(128, 256, 32), (256, 256, 64), (300, 200, 96), (1000, 2048, 768), (12800, 2048, 768),
                                   (129, 64, 40), (128, 513, 512), (4096, 200, 512)])
def test_gemm_tc_matches_fp64(M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    scale = np.abs(ref).max()
    c_tc = gemm(a, b, bias, True)
    c_ff = gemm(a, b, bias, False)
    e_tc = np.abs(c_tc - ref).max() / scale
    e_ff = np.abs(c_ff - ref).max() / scale
    print(f'M={M} N={N} K={K}: rel err tcgen05 split-fp16 {e_tc:.2e}, fp32 FFMA {e_ff:.2e}')
    assert e_ff < 2e-6
    assert e_tc < 3e-6, 'the split-precision tensor-core GEMM must stay fp32-grade'


def test_gemm_tc_special_values_and_no_bias():
    a = np.zeros((256, 64), np.float32)
    a[:, 0] = 1.0
    a[5, :] = np.linspace(-3, 3, 64, dtype=np.float32)
    b = np.eye(64, dtype=np.float32)[np.arange(128) % 64]
    c = gemm(a, b, None, True)
    ref = a @ b.T                                  # single-term sums: only the rounding of the second fp16 plane (2^-22) remains
    assert np.abs(c - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(c[:5], ref[:5])          # 0/1 data is exact


@pytest.mark.parametrize('M,N,K', [(1000, 2048, 768), (600, 200, 96), (129 + 512, 513, 64), (12800, 2048, 768)])
def test_gemm_tc_weight_multicast_clusters(M, N, K, monkeypatch):
    """KB_GEMM_MC=1: pairs of vertically adjacent tiles on 2-CTA clusters, each CTA multicasting half of the weight tile (odd tile counts
    included: the partner CTA then works on an all-padding tile).  Same MMAs in the same order as the one-CTA kernel: bit-identical."""
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = gemm(a, b, bias, True)
    monkeypatch.setenv('KB_GEMM_MC', '1')
    got = gemm(a, b, bias, True)
    assert np.array_equal(got, ref)
