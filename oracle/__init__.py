"""CPU oracle (test infrastructure, NOT product code): ctypes wrapper around gmrf_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. See the header of gmrf_oracle.c for what it restates and how it is pinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRCS = ["gmrf_oracle.c", "supernodal_cpu.c"]

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_c_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, s) for s in _SRCS if os.path.exists(os.path.join(_HERE, s))]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, *srcs, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_symbolic.restype = _c_i64
        _lib.oracle_symbolic.argtypes = [_c_i64, _i64p, _i64p, ctypes.c_void_p, _i64p, _i64p]
        _lib.oracle_factor.restype = ctypes.c_int
        _lib.oracle_factor.argtypes = [_c_i64, _i64p, _i64p, _f64p, ctypes.c_void_p, _i64p, _i64p, _f64p]
        _lib.oracle_logdet.restype = ctypes.c_double
        _lib.oracle_logdet.argtypes = [_c_i64, _i64p, _f64p]
        _lib.oracle_solve.restype = None
        _lib.oracle_solve.argtypes = [_c_i64, _i64p, _i64p, _f64p, ctypes.c_void_p, _f64p, _f64p, _c_i64]
        _lib.oracle_ltsolve.restype = None
        _lib.oracle_ltsolve.argtypes = [_c_i64, _i64p, _i64p, _f64p, ctypes.c_void_p, _f64p, _f64p, _c_i64]
        _lib.oracle_selinv.restype = None
        _lib.oracle_selinv.argtypes = [_c_i64, _i64p, _i64p, _f64p, _f64p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleFactor:
    """Simplicial LL' of P Q P' with a fixed symbolic analysis; mirrors what the reference's
    CHOLMODBackend holds (src/workspace/backend.jl:51-61)."""

    def __init__(self, Q, perm=None):
        Q = Q.tocsc()
        Q.sort_indices()
        self.n = Q.shape[0]
        self.Ap = Q.indptr.astype(np.int64)
        self.Ai = Q.indices.astype(np.int64)
        self.perm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int64)
        self.parent = np.empty(self.n, dtype=np.int64)
        self.colcount = np.empty(self.n, dtype=np.int64)
        self.nnzL = int(lib().oracle_symbolic(self.n, self.Ap, self.Ai, _ptr(self.perm), self.parent, self.colcount))
        self.Lp = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(self.colcount, out=self.Lp[1:])
        self.Li = np.zeros(self.nnzL, dtype=np.int64)
        self.Lx = np.zeros(self.nnzL, dtype=np.float64)
        self.status = None
        self.refactorize(np.ascontiguousarray(Q.data, dtype=np.float64))

    def refactorize(self, nzval):
        nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        if nzval.size != self.Ai.size:
            raise ValueError("nzval length does not match the pattern")
        self.status = int(lib().oracle_factor(self.n, self.Ap, self.Ai, nzval, _ptr(self.perm), self.Lp, self.Li, self.Lx))
        self._Z = None
        return self.status

    def logdet(self) -> float:
        return float(lib().oracle_logdet(self.n, self.Lp, self.Lx))

    def solve(self, b):
        b = np.asarray(b, dtype=np.float64)
        Bf = np.ascontiguousarray(b.reshape(self.n, -1).T).ravel()   # column-major, ld = n
        Xf = np.empty_like(Bf)
        lib().oracle_solve(self.n, self.Lp, self.Li, self.Lx, _ptr(self.perm), Bf, Xf, Bf.size // self.n)
        return Xf.reshape(-1, self.n).T.reshape(b.shape).copy()

    def backward_solve(self, z):
        z = np.asarray(z, dtype=np.float64)
        Zf = np.ascontiguousarray(z.reshape(self.n, -1).T).ravel()
        Xf = np.empty_like(Zf)
        lib().oracle_ltsolve(self.n, self.Lp, self.Li, self.Lx, _ptr(self.perm), Zf, Xf, Zf.size // self.n)
        return Xf.reshape(-1, self.n).T.reshape(z.shape).copy()

    def selinv_perm(self):
        """Z values on L's pattern, permuted ordering."""
        if self._Z is None:
            self._Z = np.empty(self.nnzL)
            lib().oracle_selinv(self.n, self.Lp, self.Li, self.Lx, self._Z)
        return self._Z

    def selinv_diag(self):
        Z = self.selinv_perm()
        d = Z[self.Lp[:-1]]
        out = np.empty(self.n)
        out[self.perm if self.perm is not None else np.arange(self.n)] = d
        return out

    def selinv(self):
        """Full symmetric CSC on the factor's pattern, original ordering (get_selinv, backend.jl:238-246)."""
        import scipy.sparse as sp
        Z = self.selinv_perm()
        p = self.perm if self.perm is not None else np.arange(self.n)
        cols = np.repeat(np.arange(self.n), np.diff(self.Lp))
        r, c = p[self.Li], p[cols]
        off = self.Li != cols
        rows = np.concatenate([r, c[off]])
        colsf = np.concatenate([c, r[off]])
        vals = np.concatenate([Z, Z[off]])
        S = sp.csc_matrix((vals, (rows, colsf)), shape=(self.n, self.n))
        S.sort_indices()
        return S
