"""GPU unit tests of the dense FP64 building blocks through the C-ABI test hooks (host pointers in/out)."""
import ctypes

import numpy as np
import pytest

from gmrf_b200 import _lib
from gmrf_b200._lib import ptr

pytestmark = pytest.mark.gpu

LOWER, ALPHA_POS, ADD_I, NAIVE, LARGE = 1, 2, 4, 8, 16


def _gemm(ta, tb, flags, m, n, k, beta, rng, lda_pad=0, ldb_pad=0, ldc_pad=0, offset=0):
    """C = (beta? C:0) + alpha * Aop Bop^T ; returns (got, want)."""
    L = _lib.lib()
    Aop = rng.standard_normal((m, k))
    Bop = rng.standard_normal((n, k))
    C0 = rng.standard_normal((m, n))
    A = np.asfortranarray(Aop.T if ta else Aop)
    B = np.asfortranarray(Bop.T if tb else Bop)
    lda = A.shape[0] + lda_pad
    ldb = B.shape[0] + ldb_pad
    ldc = m + ldc_pad
    Ab = np.zeros((lda, A.shape[1]), order="F"); Ab[: A.shape[0]] = A
    Bb = np.zeros((ldb, B.shape[1]), order="F"); Bb[: B.shape[0]] = B
    Cb = np.full((ldc, n), 7.0, order="F"); Cb[:m] = C0
    rc = L.gmrf_b200_test_gemm(0, ta, tb, flags, m, n, k, ptr(Ab), lda, ptr(Bb), ldb, float(beta), ptr(Cb), ldc)
    assert rc == 0, L.gmrf_b200_last_error(None)
    alpha = 1.0 if flags & ALPHA_POS else -1.0
    want = (C0 if beta else 0.0) + alpha * Aop @ Bop.T
    if flags & ADD_I:
        want = want + np.eye(m, n)
    got = Cb[:m].copy()
    if flags & LOWER:
        iu = np.triu_indices(m, 1, n)
        want = np.asarray(want).copy()
        want[iu] = C0[iu]
    assert np.all(Cb[m:] == 7.0)   # padding rows untouched
    return got, want


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("size", [LARGE, 0, NAIVE])
@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 128, 16), (200, 136, 77), (1, 1, 1), (5, 3, 0), (257, 130, 33),
                                   (300, 300, 300), (63, 65, 17)])
def test_gemm_variants(ta, tb, size, shape):
    rng = np.random.default_rng(hash((ta, tb, size, shape)) % 2**32)
    m, n, k = shape
    for beta, flags, pads in [(1.0, size, (0, 0, 0)), (0.0, size | ALPHA_POS, (1, 3, 1)), (1.0, size | LOWER, (2, 0, 5)),
                              (0.0, size | ADD_I, (1, 1, 1))]:
        got, want = _gemm(ta, tb, flags, m, n, k, beta, rng, *pads)
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got - want).max() <= 1e-12 * scale * max(k, 1), (shape, flags, pads)


@pytest.fixture(params=[1, 0], ids=["potrf_lookahead", "potrf_round1"])
def potrf_variant(request):
    """Both 64-column diagonal-block kernels: the blocked look-ahead one (default) and round 1's."""
    _lib.set_option("potrf_lookahead", request.param)
    yield request.param
    _lib.set_option("potrf_lookahead", 1)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 31, 32, 33, 47, 63, 64])
def test_potrf_diag(n, potrf_variant):
    L = _lib.lib()
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = M @ M.T + n * np.eye(n)
    lda = n + 3
    Ab = np.zeros((lda, n), order="F"); Ab[:n] = A
    info = ctypes.c_int(-1)
    inv = np.full((n, n), np.nan, order="F")
    assert L.gmrf_b200_test_potrf_inv(0, n, ptr(Ab), lda, ptr(inv), ctypes.byref(info)) == 0
    assert info.value == 0
    got = np.tril(Ab[:n])
    want = np.linalg.cholesky(A)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max() * n
    # the same kernel also returns inv(L) (dense n x n, zeros above the diagonal): used for TRSM-by-GEMM and the solves
    assert np.array_equal(np.triu(inv, 1), np.zeros((n, n)))
    assert np.abs(inv @ want - np.eye(n)).max() <= 1e-13 * n * np.linalg.cond(want)
    # the product's TRSM: rows below the block are solved as a GEMM with the inverted block, X = B * inv(L)^T
    m = 37
    Bm = np.asfortranarray(rng.standard_normal((m, n)))
    X = np.zeros((m, n), order="F")
    invc = np.asfortranarray(inv)
    assert L.gmrf_b200_test_gemm(0, 0, 0, 2, m, n, n, ptr(Bm), m, ptr(invc), n, 0.0, ptr(X), m) == 0
    assert np.abs(X @ want.T - Bm).max() <= 1e-12 * n * np.linalg.cond(want) * np.abs(Bm).max()
    # not positive definite: the failing column (1-based) is reported
    if n >= 3:
        A2 = A.copy(); A2[2, 2] = -1.0
        Ab = np.zeros((lda, n), order="F"); Ab[:n] = A2
        assert L.gmrf_b200_test_potrf(0, n, ptr(Ab), lda, ctypes.byref(info)) == 0
        assert info.value == 3


def test_pivot_rsqrt_accuracy():
    """The 4 x 4 pivot tiles use a straight-line 1/sqrt (hardware seed + one third-order step) instead of the library call:
    it has to be as good as a correctly rounded one over the whole range a pivot can take."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 100000), 1.0 + rng.uniform(-1e-3, 1e-3, 20000), rng.uniform(0.5, 4.0, 80000),
                        np.array([1.0, 2.0, 4.0, 0.25, 3.0, 1e-300, 1e300])])
    y = np.empty_like(x)
    assert L.gmrf_b200_test_rsqrt(0, x.size, ptr(x), ptr(y)) == 0
    want = (1.0 / np.sqrt(x.astype(np.longdouble)))
    rel = np.abs((y.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel.max() <= 2.5e-16, rel.max()                      # ~ 1 ulp
    # not-a-pivot arguments stay recognisable
    bad = np.array([-1.0, 0.0, np.nan])
    out = np.empty_like(bad)
    assert L.gmrf_b200_test_rsqrt(0, bad.size, ptr(bad), ptr(out)) == 0
    assert np.isnan(out[0]) and not np.isfinite(out[1]) and np.isnan(out[2])
