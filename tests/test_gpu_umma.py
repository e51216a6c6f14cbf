"""-m gpu: tcgen05 building blocks (TMEM, SWIZZLE_128B descriptors for K-major and MN-major fp32 operands, kind::tf32 MMA):
one tf32 pass is tf32-accurate, the three-pass hi/lo split is fp32-accurate."""
import ctypes as C

import numpy as np
import pytest

from spp_rl_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("K", [32, 256, 100])
def test_umma_tile_matches_fp64(a_mn, b_mn, K):
    lib = _lib.load_library()
    rng = np.random.RandomState(K + 2 * a_mn + b_mn)
    A = rng.randn(128, K).astype(np.float32)        # logical [M, K]
    B = rng.randn(128, K).astype(np.float32)        # logical [N, K]
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a_store = np.ascontiguousarray(A.T if a_mn else A)
    b_store = np.ascontiguousarray(B.T if b_mn else B)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    for split, tol in ((0, 2e-3), (1, 2e-6)):
        out = np.zeros((128, 128), np.float32)
        f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        _lib.check(lib.spp_umma_selftest(a_mn, b_mn, K, split, f(a_store), f(b_store), f(out)))
        err = np.abs(out - ref).max() / scale.max()
        assert err < tol, (a_mn, b_mn, K, split, err)


@pytest.mark.parametrize("a_km,b_km", [(1, 0), (0, 0), (1, 1), (0, 1)])
@pytest.mark.parametrize("M,K", [(256, 256), (256, 16), (256, 12), (256, 24), (128, 256), (100, 256), (64, 48), (4, 256), (256, 100), (256, 64)])
def test_umma_gemm256_matches_fp64(a_km, b_km, M, K):
    """the pipelined tensor-core GEMM of the fused kernels: every operand form, row tails, contraction tails"""
    lib = _lib.load_library()
    rng = np.random.RandomState(M * 7 + K + 2 * a_km + b_km)
    A = rng.randn(M, K).astype(np.float32)          # logical [M, K]
    B = rng.randn(256, K).astype(np.float32)        # logical [N, K]
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    a_store = np.ascontiguousarray(A if a_km else A.T)
    b_store = np.ascontiguousarray(B if b_km else B.T)
    out = np.zeros((M, 256), np.float32)
    f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    for reps in (1, 3):
        out[:] = 0
        _lib.check(lib.spp_umma_gemm_selftest(a_km, b_km, M, K, reps, f(a_store), f(b_store), f(out), None))
        err = np.abs(out - ref).max() / scale.max()
        assert err < 2e-6, (a_km, b_km, M, K, reps, err)
