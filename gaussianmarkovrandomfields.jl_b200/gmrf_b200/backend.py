"""Host-side mirror of the reference's `WorkspaceBackend` protocol over the C-ABI.

Reference: src/workspace/backend.jl:8-30 (protocol), :51-61 / :147-284 (CHOLMODBackend, whose caching
contract the tests pin), src/workspace/cliquetrees_backend.jl:21-150 (second implementation).
Method names, argument meaning and error behaviour follow the reference (`refactorize!` -> `refactorize`,
ArgumentError -> ValueError). Julia is not available in this image, so this Python class plays the role of
the `B200Backend <: WorkspaceBackend` glue in julia/B200Backend.jl; both bind the same entry points.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import ptr

__all__ = ["B200Backend", "PinDenseColumns", "ordering_permutation", "B200Error", "NotPositiveDefinite"]


class B200Error(RuntimeError):
    pass


class NotPositiveDefinite(B200Error):
    def __init__(self, column):
        super().__init__(f"matrix is not positive definite: non-positive pivot at column {column} of the factor")
        self.column = column


class PinDenseColumns:
    """Ordering wrapper: columns with more than frac*n nonzeros are pinned to the END of the elimination
    order, `inner` orders the remaining sparse block (src/workspace/backend.jl:63-77)."""

    def __init__(self, inner="nd", frac: float = 0.5):
        self.inner = inner
        self.frac = frac


_ORDER_CODES = {"natural": _lib.ORDER_NATURAL, "nd": _lib.ORDER_ND, "metis": _lib.ORDER_ND, "amd": _lib.ORDER_AMD,
                "mmd": _lib.ORDER_AMD}


def _csc(Q):
    Q = sp.csc_matrix(Q)
    if not Q.has_sorted_indices:
        Q = Q.copy()
        Q.sort_indices()
    return Q


def ordering_permutation(A, ordering) -> np.ndarray:
    """Resolve an ordering spec to an explicit 0-based permutation (ordering_permutation, backend.jl:93-133).
    `ordering` is a permutation vector, one of "nd"/"amd"/"natural", or a PinDenseColumns wrapper."""
    A = _csc(A)
    n = A.shape[1]
    if isinstance(ordering, PinDenseColumns):
        nnz_col = np.diff(A.indptr)
        dense = np.flatnonzero(nnz_col > ordering.frac * n)
        if dense.size == 0:
            return ordering_permutation(A, ordering.inner)
        keep = np.flatnonzero(nnz_col <= ordering.frac * n)
        sub = A[keep][:, keep]
        inner = ordering_permutation(sub, ordering.inner)
        return np.concatenate([keep[inner], dense]).astype(np.int64)
    if isinstance(ordering, str):
        code = _ORDER_CODES[ordering.lower()]
        h = _Handle(A.shape[0], A.indptr, A.indices, None, code, device=-1)
        try:
            return h.perm()
        finally:
            h.close()
    perm = np.asarray(ordering, dtype=np.int64)
    if perm.shape != (n,) or not np.array_equal(np.sort(perm), np.arange(n)):
        raise ValueError("ordering is not a permutation of 0..n-1")
    return perm


# Index base handed to the C-ABI. Python / C callers are 0-based; the Julia glue (julia/B200Backend.jl:59-66,140,156,168,
# 179) passes `SparseMatrixCSC{Float64,Int}` arrays and permutations as they are, i.e. index_base = 1. With INDEX_BASE = 1
# this binding shifts every index array on the way in and out exactly like that, so the whole test battery can exercise
# the entry path the Julia host will take (tests/test_gpu_index_base.py; env GMRF_B200_INDEX_BASE=1 flips it globally).
INDEX_BASE = int(os.environ.get("GMRF_B200_INDEX_BASE", "0"))


class index_base:
    """Context manager: `with index_base(1): ...` runs the enclosed binding calls through the 1-based C-ABI path."""

    def __init__(self, base: int):
        if base not in (0, 1):
            raise ValueError("index base must be 0 or 1")
        self.base = base

    def __enter__(self):
        global INDEX_BASE
        self.prev, INDEX_BASE = INDEX_BASE, self.base
        return self

    def __exit__(self, *exc):
        global INDEX_BASE
        INDEX_BASE = self.prev
        return False


class _Handle:
    """Owns one gmrf_b200_handle*."""

    def __init__(self, n, colptr, rowval, perm, ordering_code, device, analysis: bytes | None = None):
        L = _lib.lib()
        self._L = L
        self._h = ctypes.c_void_p()
        base = INDEX_BASE
        cp = np.ascontiguousarray(colptr, dtype=np.int64) + base
        rv = np.ascontiguousarray(rowval, dtype=np.int64) + base
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int64) + base
        if analysis is not None:       # symbolic analysis read from an exported stream instead of recomputed
            blob = np.frombuffer(analysis, dtype=np.uint8)
            rc = L.gmrf_b200_create_from_analysis(ctypes.byref(self._h), int(n), ptr(cp), ptr(rv), base, ptr(blob), blob.size, int(device))
        else:
            rc = L.gmrf_b200_create(ctypes.byref(self._h), int(n), ptr(cp), ptr(rv), base, ptr(pm), int(ordering_code), int(device))
        if rc != 0:
            msg = L.gmrf_b200_last_error(None).decode()
            self._h = None
            if rc == -1:
                raise ValueError(msg)
            raise B200Error(f"gmrf_b200_create failed ({rc}): {msg}")
        self.n = int(n)

    def check(self, rc, allow_positive=False):
        if rc == 0 or (allow_positive and rc > 0):
            return rc
        msg = self._L.gmrf_b200_last_error(self._h).decode()
        if rc == -1:
            raise ValueError(msg)
        if rc > 0:
            raise NotPositiveDefinite(rc)
        raise B200Error(f"libgmrf_b200 error {rc}: {msg}")

    def info(self) -> dict:
        v = np.zeros(len(_lib.INFO_KEYS), dtype=np.int64)
        self.check(self._L.gmrf_b200_info(self._h, ptr(v), v.size))
        return dict(zip(_lib.INFO_KEYS, (int(x) for x in v)))

    def perm(self) -> np.ndarray:
        p = np.empty(self.n, dtype=np.int64)
        self.check(self._L.gmrf_b200_get_perm(self._h, ptr(p), INDEX_BASE))
        return p - INDEX_BASE

    def export_analysis(self) -> bytes:
        """The symbolic analysis as a byte stream for `_Handle(..., analysis=...)` / `B200Backend(Q, analysis=...)`."""
        nbytes = ctypes.c_int64()
        self.check(self._L.gmrf_b200_analysis_export(self._h, None, 0, ctypes.byref(nbytes)))
        buf = np.empty(nbytes.value, dtype=np.uint8)
        self.check(self._L.gmrf_b200_analysis_export(self._h, ptr(buf), buf.size, ctypes.byref(nbytes)))
        return buf.tobytes()

    def factor_pattern(self):
        """(colptr, rowval) of the square root P'L as CSC, 0-based (symbolic: valid on analysis-only handles)."""
        nnz = ctypes.c_int64()
        self.check(self._L.gmrf_b200_factor_nnz(self._h, ctypes.byref(nnz)))
        cp = np.empty(self.n + 1, dtype=np.int64)
        rv = np.empty(nnz.value, dtype=np.int64)
        self.check(self._L.gmrf_b200_factor_pattern(self._h, ptr(cp), ptr(rv), INDEX_BASE))
        return cp - INDEX_BASE, rv - INDEX_BASE

    def close(self):
        if self._h is not None and self._h.value:
            self._L.gmrf_b200_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200Backend:
    """B200 sparse-Cholesky backend. `B200Backend(Q; ordering=None, device=0)` performs the symbolic analysis
    once and the first numeric factorization, like `CHOLMODBackend(Q::Symmetric; ordering)` (backend.jl:147-153).

    Fields mirrored from the reference struct (backend.jl:51-61) because its tests peek at them
    (test_gmrf_workspace.jl:214,219): `selinv_cache`, `selinv_diag_cache`.
    """

    def __init__(self, Q, ordering=None, device: int = 0, check: bool = False, factorize: bool = True,
                 analysis: bytes | None = None):
        Q = _csc(Q)
        if Q.shape[0] != Q.shape[1]:
            raise ValueError("Q must be square")
        self.n = Q.shape[0]
        self.check_pd = check
        perm, code = None, _lib.ORDER_ND
        if ordering is not None and analysis is None:
            if isinstance(ordering, str) and not isinstance(ordering, PinDenseColumns):
                code = _ORDER_CODES[ordering.lower()]
            else:
                perm = ordering_permutation(Q, ordering)
        self._colptr = Q.indptr.astype(np.int64)
        self._rowval = Q.indices.astype(np.int64)
        self._hd = _Handle(self.n, self._colptr, self._rowval, perm, code, device, analysis=analysis)
        self._L = self._hd._L
        self.device = device
        self.selinv_cache = None
        self.selinv_diag_cache = None
        self._selinv_pattern = None
        self._factor_pattern = None
        self._nbasis = 0
        self._pinned = []
        self.status = 0
        if factorize and device >= 0:
            self.refactorize(Q)

    # -- protocol ------------------------------------------------------------------------------------
    def refactorize(self, Q):
        """refactorize!(b, Q::Symmetric) (backend.jl:178-189): values only, pattern must be unchanged."""
        nz = Q.data if sp.issparse(Q) else np.asarray(Q)
        nz = np.ascontiguousarray(nz, dtype=np.float64)
        rc = self._L.gmrf_b200_refactorize(self._hd._h, ptr(nz), nz.size)
        self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None
        return None

    def backend_solve(self, rhs):
        """backend_solve(b, rhs) (backend.jl:191-209): new array, original ordering; vector or matrix."""
        rhs = np.asarray(rhs, dtype=np.float64)
        if rhs.shape[0] != self.n:
            raise ValueError("right-hand side has the wrong number of rows")
        B = np.asfortranarray(rhs.reshape(self.n, -1))
        X = np.empty_like(B, order="F")
        self._hd.check(self._L.gmrf_b200_solve(self._hd._h, ptr(B), ptr(X), max(self.n, 1), B.shape[1]))
        return X.reshape(rhs.shape) if rhs.ndim == 1 else X

    def backend_backward_solve(self, x):
        """backend_backward_solve(b, x) = factor.UP \\ x (backend.jl:281-284)."""
        x = np.asarray(x, dtype=np.float64)
        if x.shape[0] != self.n:
            raise ValueError("vector has the wrong length")
        Z = np.asfortranarray(x.reshape(self.n, -1))
        X = np.empty_like(Z, order="F")
        self._hd.check(self._L.gmrf_b200_solve_Lt(self._hd._h, ptr(Z), ptr(X), max(self.n, 1), Z.shape[1]))
        return X.reshape(x.shape) if x.ndim == 1 else X

    def compute_logdet(self) -> float:
        out = ctypes.c_double()
        self._hd.check(self._L.gmrf_b200_logdet(self._hd._h, ctypes.byref(out)))
        return float(out.value)

    def compute_selinv(self):
        """compute_selinv!(b) is lazy in the reference (backend.jl:215-221); the getters trigger the recursion."""
        return None

    def get_selinv(self):
        """Full symmetric CSC on the factor's pattern, original ordering, cached (backend.jl:238-246)."""
        if self.selinv_cache is None:
            if self._selinv_pattern is None:
                nnz = ctypes.c_int64()
                self._hd.check(self._L.gmrf_b200_selinv_nnz(self._hd._h, ctypes.byref(nnz)))
                cp = np.empty(self.n + 1, dtype=np.int64)
                rv = np.empty(nnz.value, dtype=np.int64)
                self._hd.check(self._L.gmrf_b200_selinv_pattern(self._hd._h, ptr(cp), ptr(rv), INDEX_BASE))
                self._selinv_pattern = (cp - INDEX_BASE, rv - INDEX_BASE)
            cp, rv = self._selinv_pattern
            vals = np.empty(rv.size, dtype=np.float64)
            self._hd.check(self._L.gmrf_b200_selinv_values(self._hd._h, ptr(vals)))
            self.selinv_cache = sp.csc_matrix((vals, rv, cp), shape=(self.n, self.n))
        return self.selinv_cache

    def get_selinv_diag(self):
        """Diagonal of Q^-1, cached until the next refactorize (backend.jl:248-257)."""
        if self.selinv_diag_cache is None:
            if self.selinv_cache is not None:
                self.selinv_diag_cache = self.selinv_cache.diagonal()
            else:
                d = np.empty(self.n, dtype=np.float64)
                self._hd.check(self._L.gmrf_b200_selinv_diag(self._hd._h, ptr(d)))
                self.selinv_diag_cache = d
        return self.selinv_diag_cache

    def selinv_extract_at(self, B):
        """Sigma read at B's pattern (backend.jl:275-279), without materialising the full selected inverse."""
        B = _csc(B)
        if B.shape != (self.n, self.n):
            raise ValueError("pattern matrix has the wrong shape")
        cp = B.indptr.astype(np.int64) + INDEX_BASE
        rv = B.indices.astype(np.int64) + INDEX_BASE
        out = np.empty(rv.size, dtype=np.float64)
        self._hd.check(self._L.gmrf_b200_selinv_extract(self._hd._h, self.n, ptr(cp), ptr(rv), INDEX_BASE, ptr(out)))
        return sp.csc_matrix((out, B.indices.copy(), B.indptr.copy()), shape=B.shape)

    def selinv_dot(self, B) -> float:
        """tr(Q^-1 B) (backend.jl:265-267): Sigma is gathered at B's pattern and contracted on the device
        (gmrf_b200_selinv_dot, fixed-shape reduction); positions outside the factor's pattern count 0."""
        B = _csc(B)
        if B.shape != (self.n, self.n):
            raise ValueError("pattern matrix has the wrong shape")
        cp = B.indptr.astype(np.int64) + INDEX_BASE
        rv = B.indices.astype(np.int64) + INDEX_BASE
        vals = np.ascontiguousarray(B.data, dtype=np.float64)
        out = ctypes.c_double()
        self._hd.check(self._L.gmrf_b200_selinv_dot(self._hd._h, self.n, ptr(cp), ptr(rv), INDEX_BASE, ptr(vals), ctypes.byref(out)))
        return float(out.value)

    def selinv_dot_basis(self) -> np.ndarray:
        """tr(Q^-1 B_j) for every array of the resident value basis: d logdet Q / d c_j for Q = sum_j c_j B_j
        (the contraction the logdetcov / logpdf pullbacks of src/workspace/autodiff.jl:8-91 need for
        fixed-pattern hyperparameter models), computed without any pattern or value upload."""
        nb = self._nbasis
        if not nb:
            raise RuntimeError("selinv_dot_basis: call set_value_basis first")
        out = np.empty(nb, dtype=np.float64)
        self._hd.check(self._L.gmrf_b200_selinv_dot_basis(self._hd._h, ptr(out), nb))
        return out

    def cholesky_sqrt(self):
        """The square root R = P'L of Q as a CSC matrix (R R' = Q): `sparse_cho_sqrt(cho)` =
        `sparse(cho.L)[invperm(cho.p), :]`, src/linear_maps/cholesky_sqrt.jl:6-21 (the matrix behind `CholeskySqrt`).
        The pattern is fetched once per backend, the values are gathered out of the factor panels on the device."""
        if self._factor_pattern is None:
            self._factor_pattern = self._hd.factor_pattern()
        cp, rv = self._factor_pattern
        vals = np.empty(rv.size, dtype=np.float64)
        self._hd.check(self._L.gmrf_b200_factor_values(self._hd._h, ptr(vals)))
        return sp.csc_matrix((vals, rv, cp), shape=(self.n, self.n))

    def export_analysis(self) -> bytes:
        """Byte stream of this backend's symbolic analysis: `B200Backend(Q2, analysis=blob)` on the same pattern skips
        ordering, elimination tree, supernodes and schedule construction (another session, another GPU of a pool)."""
        return self._hd.export_analysis()

    # -- extras --------------------------------------------------------------------------------------
    def info(self) -> dict:
        return self._hd.info()

    def permutation(self) -> np.ndarray:
        return self._hd.perm()

    def colcounts(self) -> np.ndarray:
        cc = np.empty(self.n, dtype=np.int64)
        self._hd.check(self._L.gmrf_b200_get_colcounts(self._hd._h, ptr(cc)))
        return cc

    def timings(self) -> dict:
        t = np.zeros(5)
        self._hd.check(self._L.gmrf_b200_last_timings(self._hd._h, ptr(t), 5))
        return {"h2d_ms": t[0], "factor_ms": t[1], "solve_ms": t[2], "selinv_ms": t[3], "analysis_ms": t[4]}

    def refactorize_device(self, dptr: int, nnz: int):
        rc = self._L.gmrf_b200_refactorize_device(self._hd._h, ctypes.c_void_p(dptr), int(nnz))
        self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None

    def set_value_basis(self, basis: np.ndarray):
        """Upload value arrays (nbasis x nnz, on this backend's pattern) for device-side assembly of nzval."""
        basis = np.ascontiguousarray(basis, dtype=np.float64)
        if basis.ndim != 2:
            raise ValueError("basis must be (nbasis, nnz)")
        if basis.shape[1] != self._rowval.size:
            raise ValueError(f"basis rows hold {basis.shape[1]} values but the pattern has {self._rowval.size} nonzeros")
        self._hd.check(self._L.gmrf_b200_set_value_basis(self._hd._h, ptr(basis), basis.shape[0]))
        self._nbasis = basis.shape[0]

    def refactorize_combination(self, coeff):
        """refactorize with nzval = coeff @ basis formed in HBM (no upload of nzval)."""
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        rc = self._L.gmrf_b200_refactorize_combination(self._hd._h, ptr(coeff), coeff.size)
        self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None

    def set_base_values(self, nzval):
        """Keep a prior's nzval resident in HBM (Newton loops with a diagonal observation Hessian)."""
        nz = np.ascontiguousarray(nzval, dtype=np.float64)
        if nz.size != self._rowval.size:
            raise ValueError(f"nzval holds {nz.size} values but the pattern has {self._rowval.size} nonzeros")
        self._hd.check(self._L.gmrf_b200_set_base_values(self._hd._h, ptr(nz), nz.size))

    def refactorize_minus_diag(self, hdiag):
        """refactorize Q_prior - Diagonal(hdiag) from the resident prior values: an iterate moves n doubles."""
        d = np.ascontiguousarray(hdiag, dtype=np.float64)
        rc = self._L.gmrf_b200_refactorize_base_minus_diag(self._hd._h, ptr(d), d.size)
        self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None

    def set_hessian_pattern(self, nzpos: np.ndarray):
        """nzval positions (0-based, distinct) of a sparse observation Hessian's stored entries -- `_sparse_hessian_map`
        (src/workspace/gaussian_approximation.jl:31-61), uploaded once per Newton loop."""
        p = np.ascontiguousarray(nzpos, dtype=np.int64) + INDEX_BASE
        self._hd.check(self._L.gmrf_b200_set_hessian_pattern(self._hd._h, ptr(p), p.size, INDEX_BASE))

    def refactorize_minus_sparse(self, values: np.ndarray):
        """Refactorize Q_prior - H with H given by its stored values on the pattern of `set_hessian_pattern` (iterate formed in
        HBM from the resident prior values of `set_base_values`): `_subtract_sparse_hessian!` (:74-83) + `refactorize!`."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        rc = self._L.gmrf_b200_refactorize_base_minus_sparse(self._hd._h, ptr(v), v.size)
        self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None

    def selinv_quadform_rows(self, A) -> np.ndarray:
        """diag(A Sigma A') for a sparse design matrix A (m x n): `_row_diag_AΣAt` (src/linear_predictor_marginals.jl:137-165)
        with Sigma contracted on the device; nothing but the m results comes back."""
        A = sp.csr_matrix(A, dtype=np.float64)
        A.sort_indices()
        if A.shape[1] != self.n:
            raise ValueError(f"design matrix has {A.shape[1]} columns but the field has {self.n} components")
        rp = A.indptr.astype(np.int64) + INDEX_BASE
        ci = A.indices.astype(np.int64) + INDEX_BASE
        vals = np.ascontiguousarray(A.data, dtype=np.float64)
        out = np.empty(A.shape[0], dtype=np.float64)
        self._hd.check(self._L.gmrf_b200_selinv_quadform_rows(self._hd._h, A.shape[0], ptr(rp), ptr(ci), ptr(vals), INDEX_BASE, ptr(out)))
        return out

    def lane_capacity(self) -> int:
        return int(self._L.gmrf_b200_lane_capacity(self._hd._h))

    def refactorize_lanes(self, nzvals):
        """Factorize several value sets of this pattern side by side; returns (logdets, statuses). Lane 0 stays the
        backend's factor. Capacity = _lib.set_option("lanes", B) before the backend was created."""
        nz = np.ascontiguousarray(nzvals, dtype=np.float64)
        if nz.ndim != 2 or nz.shape[1] != self._rowval.size:
            raise ValueError("nzvals must be (lanes, nnz) on this backend's pattern")
        ld = np.empty(nz.shape[0])
        st = np.zeros(nz.shape[0], dtype=np.int32)
        self._hd.check(self._L.gmrf_b200_refactorize_lanes(self._hd._h, ptr(nz), nz.shape[1], nz.shape[0], ptr(ld), ptr(st)))
        self.status = int(st[0])
        self.selinv_cache = None
        self.selinv_diag_cache = None
        return ld, st

    def refactorize_combination_lanes(self, coeffs):
        """Same with nzval = coeffs[b] @ basis formed in HBM for every lane (set_value_basis first)."""
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        if c.ndim != 2:
            raise ValueError("coeffs must be (lanes, nbasis)")
        ld = np.empty(c.shape[0])
        st = np.zeros(c.shape[0], dtype=np.int32)
        self._hd.check(self._L.gmrf_b200_refactorize_combination_lanes(self._hd._h, ptr(c), c.shape[1], c.shape[0], ptr(ld), ptr(st)))
        self.status = int(st[0])
        self.selinv_cache = None
        self.selinv_diag_cache = None
        return ld, st

    def selinv_compute(self):
        self._hd.check(self._L.gmrf_b200_selinv_compute(self._hd._h))

    def solve_device(self, dB: int, dX: int, ld: int, nrhs: int, half: bool = False):
        f = self._L.gmrf_b200_solve_Lt_device if half else self._L.gmrf_b200_solve_device
        self._hd.check(f(self._hd._h, ctypes.c_void_p(dB), ctypes.c_void_p(dX), int(ld), int(nrhs)))

    def profile_refactorize(self) -> dict:
        """Per-kernel-family device time of one refactorization (CUDA events around every launch)."""
        ms = np.zeros(4)
        cnt = np.zeros(4, dtype=np.int64)
        fl = ctypes.c_double()
        self._hd.check(self._L.gmrf_b200_profile_refactorize(self._hd._h, ptr(ms), ptr(cnt), ctypes.byref(fl)))
        names = ("gemm", "panel", "assemble", "other")
        return {"ms": dict(zip(names, ms.tolist())), "launches": dict(zip(names, cnt.tolist())), "gemm_flops": fl.value}

    def device_array(self, which: int):
        """(device pointer, doubles) of 0: panels of L, 1: inverted diagonal blocks, 2: selected-inverse panels."""
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        self._hd.check(self._L.gmrf_b200_device_array(self._hd._h, which, ctypes.byref(p), ctypes.byref(n)))
        return int(p.value or 0), int(n.value)

    def analysis_fingerprint(self) -> int:
        """64-bit fingerprint of pattern + elimination order + supernode partition + panel layout: what two handles must
        share before one can adopt the other's numeric factor."""
        f = ctypes.c_uint64()
        self._hd.check(self._L.gmrf_b200_analysis_fingerprint(self._hd._h, ctypes.byref(f)))
        return int(f.value)

    def adopt_factor(self, logdet: float, with_selinv: bool = False, fingerprint: int | None = None, status: int = 0):
        """Declare the factor arrays received from a peer GPU (sharding.broadcast_factor) to be this backend's factor.
        With the sender's `fingerprint` the library refuses a factor built on another analysis (ValueError) and the
        sender's pivot `status` becomes this backend's."""
        if fingerprint is None:
            self._hd.check(self._L.gmrf_b200_adopt_factor(self._hd._h, float(logdet), int(bool(with_selinv))))
            self.status = 0
        else:
            rc = self._L.gmrf_b200_adopt_factor_checked(self._hd._h, ctypes.c_uint64(fingerprint), float(logdet), int(status),
                                                        int(bool(with_selinv)))
            self.status = self._hd.check(rc, allow_positive=not self.check_pd)
        self.selinv_cache = None
        self.selinv_diag_cache = None

    LAUNCH_KINDS = ("assemble", "chain", "finalize", "front", "gemm_nn_s", "gemm_nn_l", "gemm_nt_s", "gemm_nt_l", "gemm_tt_s",
                    "gemm_tt_l", "gather", "transpose", "fwd_asm", "fwd_step", "bwd_gather", "bwd_step", "panel", "split_reduce", "fwd_asm_wide", "rows_gather", "bwd_reduce", "assemble_g", "fwd_wstep", "fwd_wdiag", "bwd_wstep", "bwd_wdiag")

    def profile_plan(self, phase: int, nrhs: int = 1):
        """Per-launch (kind, grid, ms) of one phase: 0 factorization, 1 selinv, 2 forward sweep, 3 backward sweep."""
        cnt = ctypes.c_int64()
        self._hd.check(self._L.gmrf_b200_profile_plan(self._hd._h, phase, nrhs, 0, None, None, None, ctypes.byref(cnt)))
        n = cnt.value
        kind = np.zeros(n, dtype=np.int32); grid = np.zeros(n, dtype=np.int32); ms = np.zeros(n)
        self._hd.check(self._L.gmrf_b200_profile_plan(self._hd._h, phase, nrhs, n, ptr(kind), ptr(grid), ptr(ms), ctypes.byref(cnt)))
        return [(self.LAUNCH_KINDS[k], int(g), float(t)) for k, g, t in zip(kind, grid, ms)]

    def plan_launch_info(self, phase: int):
        """(flops, kmax, ntasks) arrays of the launches of a phase, in `profile_plan`'s order (GEMM launches only carry flops)."""
        cnt = ctypes.c_int64()
        self._hd.check(self._L.gmrf_b200_plan_launch_info(self._hd._h, phase, 0, None, None, None, ctypes.byref(cnt)))
        n = cnt.value
        fl = np.zeros(n); km = np.zeros(n, dtype=np.int32); nt = np.zeros(n, dtype=np.int32)
        self._hd.check(self._L.gmrf_b200_plan_launch_info(self._hd._h, phase, n, ptr(fl), ptr(km), ptr(nt), ctypes.byref(cnt)))
        return fl, km, nt

    def pin_host_buffer(self, arr: np.ndarray) -> bool:
        """Page-lock a caller-owned numpy buffer for asynchronous H2D copies (no-op without a device)."""
        # Only large buffers are registered: they are mmap-backed (own pages), whereas page-locking a small heap
        # array would also pin its neighbours' pages and make later pageable copies of those fail.
        if self.device < 0 or arr.nbytes < (64 << 20):
            return False
        rc = self._L.gmrf_b200_host_register(ptr(arr), arr.nbytes)
        if rc == 0:
            self._pinned.append(arr)
        return rc == 0

    def close(self):
        for a in self._pinned:
            self._L.gmrf_b200_host_unregister(ptr(a))
        self._pinned = []
        self._hd.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
