"""CPU baseline (bench/test infrastructure, NOT product code): multithreaded BLAS-3 supernodal Cholesky + logdet on
the host cores -- the stand-in for the reference's CHOLMOD path, which cannot run in this image (see
supernodal_cpu.c). BLAS/LAPACK come from the OpenBLAS bundled with SciPy, via the function-pointer capsules of
scipy.linalg.cython_blas / cython_lapack. Only bench.py (cpu_baseline, --impl reference) and tests/ use this."""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np

from . import lib as _oracle_lib

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _capsule_ptr(capsule):
    ctypes.pythonapi.PyCapsule_GetName.restype = ctypes.c_char_p
    ctypes.pythonapi.PyCapsule_GetName.argtypes = [ctypes.py_object]
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    name = ctypes.pythonapi.PyCapsule_GetName(capsule)
    return ctypes.c_void_p(ctypes.pythonapi.PyCapsule_GetPointer(capsule, name))


def _blas_ptrs():
    import scipy.linalg.cython_blas as cb
    import scipy.linalg.cython_lapack as cl
    return (_capsule_ptr(cl.__pyx_capi__["dpotrf"]), _capsule_ptr(cb.__pyx_capi__["dtrsm"]),
            _capsule_ptr(cb.__pyx_capi__["dsyrk"]))


class CpuSupernodalCholesky:
    """Numeric refactorization on the CPU over the symbolic tables of an analysis-only library handle
    (host-side integer work is shared; every floating-point operation here runs on the host cores)."""

    def __init__(self, tables, threads: int | None = None):
        T = tables
        self.T = T
        self.threads = threads or (os.cpu_count() or 1)
        L = _oracle_lib()
        f = L.cpu_supernodal_factor_level
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int64, _i64p] + [_i64p] * 10 + [_f64p, _f64p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.cpu_scatter.restype = None
        L.cpu_scatter.argtypes = [ctypes.c_int64, _f64p, ctypes.c_int64, _i64p, _i64p, _f64p]
        L.cpu_logdet.restype = ctypes.c_double
        L.cpu_logdet.argtypes = [ctypes.c_int64, _i64p, _f64p]
        self._L = L
        ns = T.nsuper
        counts = np.zeros(ns + 1, dtype=np.int64)
        for s in range(ns):
            p = T.sparent[s]
            if p >= 0:
                counts[p + 1] += 1
        self.child_ptr = np.cumsum(counts)
        self.child_idx = np.zeros(max(int(self.child_ptr[-1]), 1), dtype=np.int64)
        w = self.child_ptr[:-1].copy()
        for s in range(ns):
            p = T.sparent[s]
            if p >= 0:
                self.child_idx[w[p]] = s
                w[p] += 1
        order = np.argsort(T.level, kind="stable")
        self.levels = [np.ascontiguousarray(order[T.level[order] == l]) for l in range(int(T.level.max()) + 1 if ns else 0)]
        nr = (T.row_ptr[1:] - T.row_ptr[:-1]) - (T.super_ptr[1:] - T.super_ptr[:-1])
        self.upd = np.zeros(max(int((T.upd_off + T.upd_ld * nr).max()) if ns else 1, 1))
        self.Lx = np.zeros(int(T.panel_off[-1]))
        self.diag_pos = np.concatenate([
            T.panel_off[s] + np.arange(T.ns(s)) * (T.panel_ld[s] + 1) for s in range(ns)]).astype(np.int64) if ns else np.zeros(0, np.int64)
        self.ptrs = _blas_ptrs()

    def _blas_threads(self, k: int):
        """Context limiting the BLAS pool to k threads. The controller is built once: discovering the loaded BLAS libraries
        costs ~2 ms, which per level of a sweep would be a large share of a CPU solve."""
        if getattr(self, "_tpc", None) is None:
            from threadpoolctl import ThreadpoolController
            self._tpc = ThreadpoolController()
        return self._tpc.limit(limits=k, user_api="blas")

    def refactorize(self, nzval) -> float:
        """Returns the wall time in seconds of scatter + numeric factorization + logdet."""
        T = self.T
        nz = np.ascontiguousarray(nzval, dtype=np.float64)
        t0 = time.perf_counter()
        self._L.cpu_scatter(self.Lx.size, self.Lx, T.q_src.size, T.q_src, T.q_dst, nz)
        status = 0
        for sup in self.levels:
            parallel = 1 if sup.size >= 2 * self.threads else 0
            with self._blas_threads(1 if parallel else self.threads):
                r = self._L.cpu_supernodal_factor_level(
                    sup.size, sup, T.super_ptr, T.row_ptr, T.row_idx, T.rel_idx, T.panel_off, T.panel_ld,
                    T.upd_off, T.upd_ld, self.child_ptr, self.child_idx, self.Lx, self.upd, *self.ptrs, parallel)
            if r and not status:
                status = r
        self.status = status
        self.logdet = float(self._L.cpu_logdet(self.diag_pos.size, self.diag_pos, self.Lx))
        return time.perf_counter() - t0

    def selinv(self) -> float:
        """Supernodal Takahashi selected inversion of the current factor on the host cores (top-down over the levels,
        OpenMP across the fronts of a level / threaded BLAS inside the big ones). Returns the wall time in seconds; the
        result stays in `self.Zx` (same panel layout as the factor), `selinv_diag()` reads the marginal variances."""
        import scipy.linalg.cython_blas as cb
        T = self.T
        L = self._L
        f = L.cpu_supernodal_selinv_level
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int64, _i64p] + [_i64p] * 7 + [_f64p, _f64p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                          ctypes.c_void_p, ctypes.c_int]
        L.cpu_supernodal_selinv_cleanup.restype = None
        L.cpu_supernodal_selinv_cleanup.argtypes = [ctypes.c_int64, ctypes.c_void_p]
        gemm = _capsule_ptr(cb.__pyx_capi__["dgemm"])
        trsm = self.ptrs[1]
        ns = T.nsuper
        if getattr(self, "Zx", None) is None:
            self.Zx = np.zeros_like(self.Lx)
        W = (ctypes.c_void_p * max(ns, 1))()
        pending = np.ascontiguousarray(np.diff(self.child_ptr), dtype=np.int32)
        sparent = np.ascontiguousarray(T.sparent, dtype=np.int64)
        t0 = time.perf_counter()
        rc = 0
        try:
            for sup in reversed(self.levels):
                parallel = 1 if sup.size >= 2 * self.threads else 0
                with self._blas_threads(1 if parallel else self.threads):
                    rc = f(sup.size, sup, T.super_ptr, sparent, T.row_ptr, T.rel_idx, T.panel_off, T.panel_ld, self.child_ptr,
                           self.Lx, self.Zx, ctypes.cast(W, ctypes.c_void_p), pending.ctypes.data_as(ctypes.c_void_p), gemm, trsm,
                           parallel)
                if rc:
                    raise MemoryError("cpu selinv: could not allocate a W block")
        finally:
            L.cpu_supernodal_selinv_cleanup(ns, ctypes.cast(W, ctypes.c_void_p))
        return time.perf_counter() - t0

    def solve(self, b, half: bool = False):
        """(x, seconds): Q x = b (or x = P' L^-T b with `half=True`, the sampling half solve) with the current factor, one
        right-hand side, supernodal forward / backward sweeps level by level on the host cores."""
        import scipy.linalg.cython_blas as cb
        T = self.T
        f = self._L.cpu_supernodal_solve_level
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int64, _i64p] + [_i64p] * 8 + [_f64p, _f64p, _f64p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        trsv, gemv = _capsule_ptr(cb.__pyx_capi__["dtrsv"]), _capsule_ptr(cb.__pyx_capi__["dgemv"])
        if getattr(self, "_u", None) is None:
            self._u = np.zeros(max(int(T.row_ptr[-1]), 1))
        b = np.asarray(b, dtype=np.float64)
        y = np.ascontiguousarray(b[T.perm]) if not half else b.copy()      # F.UP \\ x = P' L^-T x: the input is in factor coordinates
        t0 = time.perf_counter()
        sweeps = ([(self.levels, 0)] if not half else []) + [(list(reversed(self.levels)), 1)]
        with self._blas_threads(1):        # one thread pool for the whole sweep: OpenMP over fronts, or over the rows of a big one
            for levels, backward in sweeps:
                for sup in levels:
                    parallel = 1 if sup.size >= 2 else 0
                    f(sup.size, sup, T.super_ptr, T.row_ptr, T.row_idx, T.rel_idx, T.panel_off, T.panel_ld, self.child_ptr,
                      self.child_idx, self.Lx, y, self._u, trsv, gemv, backward, parallel)
        dt = time.perf_counter() - t0
        x = np.empty_like(y)
        x[T.perm] = y
        return x, dt

    def selinv_diag(self) -> np.ndarray:
        """diag(Q^-1) in the original ordering from the Z panels of the last `selinv()`."""
        d = np.empty(self.T.n)
        d[self.T.perm] = self.Zx[self.diag_pos]
        return d
