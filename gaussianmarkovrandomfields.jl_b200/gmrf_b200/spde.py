"""Synthetic inputs for the sparse-Cholesky hot path: precision matrices with the
sparsity structure and conditioning the reference hands to its factorization backend.

These are *producers of Q*, not part of the hot path.
Ferrite/Gmsh are unavailable, so structured simplicial meshes reproduce the reference's recipe:

* P1 lumped mass / stiffness: ext/GaussianMarkovRandomFieldsFEM/fem_utils.jl:6-8, 42-70, 86-110
* Matern alpha-recursion:      ext/GaussianMarkovRandomFieldsFEM/matern_spde.jl:177-231, 332-356
* nu / alpha / kappa:          matern_spde.jl:343, 415-422
* structural pattern S^alpha:  matern_spde.jl:248-265, fem_utils.jl:313-335
* deterministic test fixtures: test/workspace/test_backend_ordering.jl:9-17,
                               benchmarks/benchmarks.jl:160-179,
                               test/workspace/test_workspace_gaussian_approximation.jl:8-32

Everything returns scipy CSC matrices holding the FULL symmetric pattern with sorted row
indices, matching `SparseMatrixCSC{Float64,Int}` as stored in `GMRFWorkspace.Q`
(src/workspace/gmrf_workspace.jl:32).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

__all__ = [
    "grid_border_fixture", "grid3d_fixture", "tridiag_fixture", "random_spd_fixture",
    "mesh2d", "mesh3d", "p1_mass_stiffness", "matern_pattern", "MaternSPDE",
    "geometric_nd_perm",
]


# --------------------------------------------------------------------------- fixtures
def grid_border_fixture(nx: int = 12) -> sp.csc_matrix:
    """12x12 grid Laplacian + 0.1 I with a dense border row/col h=0.01 and corner 2.0
    (test/workspace/test_backend_ordering.jl:9-17). N = nx*nx + 1."""
    n = nx * nx
    a1 = sp.diags([2.0 * np.ones(nx), -np.ones(nx - 1), -np.ones(nx - 1)], [0, 1, -1])
    eye = sp.identity(nx)
    qgrid = sp.kron(eye, a1) + sp.kron(a1, eye) + 0.1 * sp.identity(n)
    h = sp.csc_matrix(np.full((n, 1), 0.01))
    q = sp.bmat([[qgrid, h], [h.T, sp.csc_matrix(np.array([[2.0]]))]], format="csc")
    q.sort_indices()
    return q


def grid3d_fixture(nx: int = 12, ny: int = 12, nz: int = 12, c: float = 0.1) -> sp.csc_matrix:
    """7-point 3D grid precision, diagonal = degree + c (benchmarks/benchmarks.jl:160-179)."""
    def path(m):
        return sp.diags([np.ones(m - 1), np.ones(m - 1)], [1, -1])
    ix, iy, iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    adj = sp.kron(sp.kron(iz, iy), path(nx)) + sp.kron(sp.kron(iz, path(ny)), ix) \
        + sp.kron(sp.kron(path(nz), iy), ix)
    adj = sp.csc_matrix(adj)
    deg = np.asarray(adj.sum(axis=1)).ravel()
    q = sp.csc_matrix(sp.diags(deg + c) - adj)
    q.sort_indices()
    return q


def tridiag_fixture(n: int = 10, diag: float = 2.0, off: float = -0.8) -> sp.csc_matrix:
    """spdiagm(0=>diag, +-1=>off) (test_workspace_gaussian_approximation.jl:8-10)."""
    q = sp.diags([diag * np.ones(n), off * np.ones(n - 1), off * np.ones(n - 1)], [0, 1, -1], format="csc")
    q.sort_indices()
    return q


def random_spd_fixture(n: int = 20, density: float = 0.3, seed: int = 42) -> sp.csc_matrix:
    """sprand-like A, Q = A A' + n I (test/workspace/test_gmrf_workspace.jl:8-13). Julia's
    MersenneTwister stream is not reproducible here; numpy's Generator(seed) is used instead."""
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=density, random_state=rng, format="csc")
    q = sp.csc_matrix(a @ a.T + n * sp.identity(n))
    q.sort_indices()
    return q


# --------------------------------------------------------------------------- meshes
def mesh2d(nx: int, ny: int | None = None):
    """Structured triangulation of [-1,1]^2: nx x ny cells, each split into 2 triangles
    (Ferrite `generate_grid(Triangle, (nx, ny))` layout). Returns (coords (nv,2), cells (ne,3))."""
    ny = nx if ny is None else ny
    xs = np.linspace(-1.0, 1.0, nx + 1)
    ys = np.linspace(-1.0, 1.0, ny + 1)
    xx, yy = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([xx.ravel(), yy.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    t1 = np.stack([v00, v10, v01], axis=1)
    t2 = np.stack([v10, v11, v01], axis=1)
    return coords, np.concatenate([t1, t2], axis=0).astype(np.int64)


def mesh3d(nx: int, ny: int | None = None, nz: int | None = None):
    """Structured tetrahedralisation of [-1,1]^3: nx x ny x nz cells, Kuhn 6-tet split.
    Returns (coords (nv,3), cells (ne,4))."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    xs, ys, zs = (np.linspace(-1.0, 1.0, m + 1) for m in (nx, ny, nz))
    zz, yy, xx = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (k * (ny + 1) * (nx + 1) + j * (nx + 1) + i).ravel()
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    tets = []
    # Kuhn: one tet per permutation of the axes, path 000 -> ... -> 111
    import itertools
    for perm in itertools.permutations((sx, sy, sz)):
        a = base
        b = a + perm[0]
        c = b + perm[1]
        d = c + perm[2]
        tets.append(np.stack([a, b, c, d], axis=1))
    return coords, np.concatenate(tets, axis=0).astype(np.int64)


def p1_mass_stiffness(coords: np.ndarray, cells: np.ndarray):
    """Row-sum-lumped P1 mass (diagonal, as a vector) and P1 stiffness G (CSC), H = I.
    fem_utils.jl:6-8 (lump), :42-70 (mass), :86-110 (diffusion)."""
    nv, d = coords.shape
    ne = cells.shape[0]
    x = coords[cells]                                  # (ne, d+1, d)
    edges = x[:, 1:, :] - x[:, :1, :]                   # (ne, d, d) rows = edge vectors
    det = np.linalg.det(edges)
    vol = np.abs(det) / math.factorial(d)
    # gradients of barycentric coordinates: rows of inv(edges) transposed
    inv = np.linalg.inv(edges)                          # (ne, d, d): inv @ edges = I
    grads = np.empty((ne, d + 1, d))
    grads[:, 1:, :] = np.transpose(inv, (0, 2, 1))
    grads[:, 0, :] = -grads[:, 1:, :].sum(axis=1)
    ge = np.einsum("eik,ejk->eij", grads, grads) * vol[:, None, None]
    rows = np.repeat(cells, d + 1, axis=1).ravel()
    cols = np.tile(cells, (1, d + 1)).ravel()
    g = sp.coo_matrix((ge.ravel(), (rows, cols)), shape=(nv, nv)).tocsc()
    g.sum_duplicates()
    g.sort_indices()
    c = np.zeros(nv)
    np.add.at(c, cells.ravel(), np.repeat(vol / (d + 1), d + 1))
    return c, g


def _pattern_power(s: sp.csc_matrix, alpha: int) -> sp.csc_matrix:
    ones = sp.csc_matrix((np.ones(s.nnz), s.indices.copy(), s.indptr.copy()), shape=s.shape)
    p = ones
    for _ in range(alpha - 1):
        p = sp.csc_matrix(p @ ones)
        p.data[:] = 1.0
    p.sort_indices()
    return p


def matern_pattern(g: sp.csc_matrix, alpha: int) -> sp.csc_matrix:
    """Structural pattern S^alpha with S = I u pattern(G) (matern_spde.jl:248-265)."""
    n = g.shape[0]
    s = sp.csc_matrix((np.ones(g.nnz), g.indices, g.indptr), shape=g.shape) + sp.identity(n, format="csc")
    s = sp.csc_matrix(s)
    s.sort_indices()
    return _pattern_power(s, alpha)


def _scatter_into_pattern(pattern: sp.csc_matrix, q: sp.csc_matrix) -> np.ndarray:
    """Values of q laid out on `pattern` (explicit zeros kept): fem_utils.jl:313-335."""
    n = pattern.shape[0]
    q = sp.csc_matrix(q)
    q.sort_indices()
    pcol = np.repeat(np.arange(n, dtype=np.int64), np.diff(pattern.indptr))
    qcol = np.repeat(np.arange(n, dtype=np.int64), np.diff(q.indptr))
    pkey = pcol * n + pattern.indices
    qkey = qcol * n + q.indices
    pos = np.searchsorted(pkey, qkey)
    if not np.array_equal(pkey[pos], qkey):
        raise ValueError("matrix has entries outside the structural pattern")
    out = np.zeros(pattern.nnz)
    out[pos] = q.data
    return out


class MaternSPDE:
    """MaternModel precision on a fixed mesh: `values(tau, range)` returns nzval on a pattern
    that does not depend on the hyperparameters (ext/.../matern_model.jl:109-121)."""

    def __init__(self, coords, cells, smoothness: int):
        self.d = coords.shape[1]
        self.n = coords.shape[0]
        self.c, self.g = p1_mass_stiffness(coords, cells)
        self.nu = smoothness + 1.0 if self.d % 2 == 0 else smoothness + 0.5
        alpha = self.nu + self.d / 2.0
        assert abs(alpha - round(alpha)) < 1e-12
        self.alpha = int(round(alpha))
        self.pattern = matern_pattern(self.g, self.alpha)
        self.colptr = self.pattern.indptr.astype(np.int64)
        self.rowval = self.pattern.indices.astype(np.int64)

    def precision(self, tau: float = 1.0, range_: float = 0.3) -> sp.csc_matrix:
        nu, d = self.nu, self.d
        kappa = math.sqrt(8.0 * nu) / range_
        ratio = math.gamma(nu) / (math.gamma(nu + d / 2.0) * (4.0 * math.pi) ** (d / 2.0) * kappa ** (2.0 * nu))
        k = sp.csc_matrix(kappa ** 2 * sp.diags(self.c) + self.g)
        cinv = sp.diags(1.0 / self.c)
        if self.alpha == 1:
            q = ratio * k
        else:
            # alpha even: start from K C^-1 K ; alpha odd: start from K
            q = k if self.alpha % 2 == 1 else k @ cinv @ k
            a = 1 if self.alpha % 2 == 1 else 2
            while a < self.alpha:
                q = k @ cinv @ q @ cinv @ k
                a += 2
            q = ratio * q
        q = sp.csc_matrix(tau * q)
        vals = _scatter_into_pattern(self.pattern, q)
        return sp.csc_matrix((vals, self.pattern.indices, self.pattern.indptr), shape=q.shape)

    def values(self, tau: float = 1.0, range_: float = 0.3) -> np.ndarray:
        return self.precision(tau, range_).data


# --------------------------------------------------------------------------- orderings
def geometric_nd_perm(dims, leaf: int = 64, width: int = 1) -> np.ndarray:
    """Geometric nested-dissection permutation of a structured vertex grid `dims`
    (x fastest). `width` = separator thickness (stencil hop count). Returns perm with
    perm[k] = original index of the k-th eliminated vertex (0-based). This is a host-side
    ordering a caller may pass as `ordering=perm` (src/workspace/backend.jl:147-153)."""
    dims = tuple(int(v) for v in dims)
    nd = len(dims)
    strides = np.cumprod((1,) + dims[:-1])
    out = []

    def rec(lo, hi):
        ext = [h - l for l, h in zip(lo, hi)]
        npts = int(np.prod(ext))
        if npts == 0:
            return
        ax = int(np.argmax(ext))
        if npts <= leaf or ext[ax] <= 2 * width:
            out.append(_box_indices(lo, hi, strides))
            return
        mid = lo[ax] + (ext[ax] - width) // 2
        lo_a, hi_a = list(lo), list(hi)
        hi_a[ax] = mid
        lo_b, hi_b = list(lo), list(hi)
        lo_b[ax] = mid + width
        lo_s, hi_s = list(lo), list(hi)
        lo_s[ax], hi_s[ax] = mid, mid + width
        rec(lo_a, hi_a)
        rec(lo_b, hi_b)
        out.append(_box_indices(lo_s, hi_s, strides))

    import sys
    sys.setrecursionlimit(10000)
    rec([0] * nd, list(dims))
    perm = np.concatenate(out).astype(np.int64)
    assert perm.size == int(np.prod(dims))
    return perm


def _box_indices(lo, hi, strides):
    grids = np.meshgrid(*[np.arange(l, h, dtype=np.int64) for l, h in zip(lo, hi)], indexing="ij")
    idx = np.zeros_like(grids[0])
    for g, s in zip(grids, strides):
        idx = idx + g * int(s)
    # x fastest inside the box
    return np.transpose(idx, tuple(reversed(range(len(lo))))).ravel()
