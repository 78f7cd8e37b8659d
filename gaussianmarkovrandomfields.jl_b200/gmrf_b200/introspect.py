"""Symbolic-schedule tables of a handle, fetched through the C-ABI introspection calls (host-side facts only)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import ptr


class Tables:
    def __init__(self, handle):
        L = _lib.lib()
        h = handle._h
        info = handle.info()
        self.info = info
        n, ns = info["n"], info["nsuper"]
        self.n, self.nsuper = n, ns
        i64 = lambda k: np.zeros(k, dtype=np.int64)
        self.super_ptr, self.sparent, self.level = i64(ns + 1), i64(ns), i64(ns)
        self.row_ptr, self.panel_off, self.panel_ld = i64(ns + 1), i64(ns + 1), i64(ns)
        self.upd_off, self.upd_ld = i64(ns), i64(ns)
        assert L.gmrf_b200_get_supernodes(h, ptr(self.super_ptr), ptr(self.sparent), ptr(self.level), ptr(self.row_ptr),
                                          ptr(self.panel_off), ptr(self.panel_ld), ptr(self.upd_off), ptr(self.upd_ld)) == 0
        tot = int(self.row_ptr[-1])
        self.row_idx, self.rel_idx = i64(tot), i64(tot)
        assert L.gmrf_b200_get_rows(h, ptr(self.row_idx), ptr(self.rel_idx)) == 0
        cnt = ctypes.c_int64()
        assert L.gmrf_b200_get_scatter(h, ctypes.byref(cnt), None, None) == 0
        self.q_src, self.q_dst = i64(cnt.value), i64(cnt.value)
        assert L.gmrf_b200_get_scatter(h, ctypes.byref(cnt), ptr(self.q_src), ptr(self.q_dst)) == 0
        self.perm = handle.perm()
        self.colcount = i64(n)
        assert L.gmrf_b200_get_colcounts(h, ptr(self.colcount)) == 0
        self.parent = i64(n)
        assert L.gmrf_b200_get_etree(h, ptr(self.parent)) == 0
        self.children = [[] for _ in range(ns)]
        for s in range(ns):
            if self.sparent[s] >= 0:
                self.children[self.sparent[s]].append(s)

    def ns(self, s): return int(self.super_ptr[s + 1] - self.super_ptr[s])
    def nrow(self, s): return int(self.row_ptr[s + 1] - self.row_ptr[s])
    def rows(self, s): return self.row_idx[self.row_ptr[s]:self.row_ptr[s + 1]]
    def rel(self, s): return self.rel_idx[self.row_ptr[s] + self.ns(s):self.row_ptr[s + 1]]

    def panel(self, Lx, s):
        ld, ns = int(self.panel_ld[s]), self.ns(s)
        off = int(self.panel_off[s])
        return Lx[off:off + ld * ns].reshape(ns, ld).T[: self.nrow(s), :]   # view, column-major
