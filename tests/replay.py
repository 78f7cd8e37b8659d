"""Host replay of the library's supernodal schedule with dense numpy blocks (TEST infrastructure).

It consumes ONLY the symbolic tables exported by the C-ABI introspection calls of an analysis-only handle and
re-executes, supernode by supernode, the exact algebra the CUDA plans implement (assemble / POTRF / TRSM / SYRK,
multifrontal forward + gathered backward solve, Takahashi with the [G; T'] L11^-1 formulation). If this replay
matches the oracle, the integer tables (row structures, relative indices, scatter map, pool offsets, levels)
and the formulas are right, and any GPU mismatch is a kernel bug.
"""
from __future__ import annotations

import ctypes

import numpy as np
import scipy.linalg as sla

from gmrf_b200 import _lib
from gmrf_b200._lib import ptr


from gmrf_b200.introspect import Tables  # noqa: E402,F401


def check_structure(T: Tables):
    """Invariants of the tables."""
    n = T.n
    assert np.array_equal(np.sort(T.perm), np.arange(n))
    assert T.super_ptr[0] == 0 and T.super_ptr[-1] == n
    live = {}
    for s in range(T.nsuper):
        rows, ns = T.rows(s), T.ns(s)
        assert np.array_equal(rows[:ns], np.arange(T.super_ptr[s], T.super_ptr[s + 1]))
        assert np.all(np.diff(rows) > 0)
        p = T.sparent[s]
        if p >= 0:
            assert p > s and T.level[p] > T.level[s]
            prow = T.rows(p)
            assert np.array_equal(prow[T.rel(s)], rows[ns:])
        else:
            assert T.nrow(s) == ns
        assert T.panel_ld[s] >= T.nrow(s)
        assert T.panel_off[s + 1] - T.panel_off[s] >= T.panel_ld[s] * ns
    # update-pool intervals alive at the same time must not overlap
    order = np.argsort(T.level, kind="stable")
    events = []
    for s in range(T.nsuper):
        nr = T.nrow(s) - T.ns(s)
        if nr == 0:
            continue
        birth = int(T.level[s])
        death = int(T.level[T.sparent[s]])
        events.append((birth, death, int(T.upd_off[s]), int(T.upd_off[s] + T.upd_ld[s] * nr)))
    events.sort()
    for i, (b, d, lo, hi) in enumerate(events[:2000]):
        for (b2, d2, lo2, hi2) in events[i + 1:i + 200]:
            if b2 > d:
                break
            assert hi <= lo2 or hi2 <= lo, "overlapping live update matrices"


def factor(T: Tables, nzval):
    """Multifrontal numeric factorization; returns the panel array (same layout as the device's)."""
    Lx = np.zeros(int(T.panel_off[-1]))
    Lx[T.q_dst] = nzval[T.q_src]
    upd = {}
    for s in range(T.nsuper):   # postorder: children before parents
        ns, nrow = T.ns(s), T.nrow(s)
        nr = nrow - ns
        P = T.panel(Lx, s)
        U = np.zeros((nr, nr))
        for c in T.children[s]:
            rel = T.rel(c)
            Uc = upd.pop(c)
            inS = rel < ns
            # lower triangle of the child's update, scattered through the relative indices
            ii, jj = np.tril_indices(len(rel))
            r, cidx = rel[ii], rel[jj]
            v = Uc[ii, jj]
            m1 = cidx < ns
            np.add.at(P, (r[m1], cidx[m1]), v[m1])
            m2 = ~m1
            np.add.at(U, (r[m2] - ns, cidx[m2] - ns), v[m2])
        L11 = np.linalg.cholesky(np.tril(P[:ns, :ns]) + np.tril(P[:ns, :ns], -1).T)
        P[:ns, :ns] = L11
        if nr:
            L21 = sla.solve_triangular(L11, P[ns:, :].T, lower=True).T
            P[ns:, :] = L21
            U -= np.tril(L21 @ L21.T)
            upd[s] = U
    return Lx


def logdet(T: Tables, Lx):
    acc = 0.0
    for s in range(T.nsuper):
        acc += np.log(np.diag(T.panel(Lx, s)[: T.ns(s), : T.ns(s)])).sum()
    return 2.0 * acc


def solve(T: Tables, Lx, b, half=False):
    """x = Q^-1 b (half=False) or x = P' L^-T z (half=True)."""
    b = np.asarray(b, dtype=float)
    y = b.copy() if half else b[T.perm].copy()
    if not half:
        u = {}
        for s in range(T.nsuper):
            ns, nrow = T.ns(s), T.nrow(s)
            f = int(T.super_ptr[s])
            P = T.panel(Lx, s)
            us = np.zeros(nrow - ns)
            for c in T.children[s]:
                rel, uc = T.rel(c), u.pop(c)
                m = rel < ns
                y[f + rel[m]] += uc[m]
                us[rel[~m] - ns] += uc[~m]
            y[f:f + ns] = sla.solve_triangular(P[:ns, :ns], y[f:f + ns], lower=True)
            us -= P[ns:, :] @ y[f:f + ns]
            u[s] = us
    for s in range(T.nsuper - 1, -1, -1):
        ns = T.ns(s)
        f = int(T.super_ptr[s])
        P = T.panel(Lx, s)
        rows = T.rows(s)
        rhs = y[f:f + ns] - P[ns:, :].T @ y[rows[ns:]]
        y[f:f + ns] = sla.solve_triangular(P[:ns, :ns], rhs, lower=True, trans="T")
    x = np.empty_like(y)
    x[T.perm] = y
    return x


def selinv(T: Tables, Lx):
    """Takahashi recursion as the device plan does it. Returns the Z panel array."""
    Zx = np.zeros_like(Lx)
    W = {}
    for s in range(T.nsuper - 1, -1, -1):
        ns, nrow = T.ns(s), T.nrow(s)
        nr = nrow - ns
        P, Z = T.panel(Lx, s), T.panel(Zx, s)
        L11, L21 = np.tril(P[:ns, :ns]), P[ns:, :]
        if nr:
            p = int(T.sparent[s])
            rel, pns = T.rel(s), T.ns(p)
            Zp, Wp = T.panel(Zx, p), W[p]
            Ws = np.empty((nr, nr))
            for b in range(nr):
                pb = rel[b]
                a = np.arange(b, nr)
                col = Zp[rel[a], pb] if pb < pns else Wp[rel[a] - pns, pb - pns]
                Ws[a, b] = col
                Ws[b, a] = col
            Tm = -Ws @ L21
        else:
            Ws = np.zeros((0, 0))
            Tm = np.zeros((0, ns))
        G = np.eye(ns) - L21.T @ Tm
        X = np.vstack([G, Tm]) @ np.linalg.inv(L11)        # [H; Z_RS]
        H = X[:ns, :]
        Zss = H.T @ np.linalg.inv(L11)
        Z[:ns, :] = Zss
        Z[ns:, :] = X[ns:, :]
        W[s] = Ws
    return Zx


def selinv_diag(T: Tables, Zx):
    d = np.empty(T.n)
    for s in range(T.nsuper):
        ns, f = T.ns(s), int(T.super_ptr[s])
        d[T.perm[f:f + ns]] = np.diag(T.panel(Zx, s)[:ns, :ns])
    return d
