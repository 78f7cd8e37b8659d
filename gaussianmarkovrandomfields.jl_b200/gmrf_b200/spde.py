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

    def __init__(self, coords, cells, smoothness: int, diffusion: float = 1.0):
        self.d = coords.shape[1]
        self.n = coords.shape[0]
        self.c, self.g = p1_mass_stiffness(coords, cells)
        if diffusion != 1.0:                      # diffusion_factor H = diffusion * I (fem_utils.jl:86-110)
            self.g = sp.csc_matrix(self.g * diffusion)
        self.nu = smoothness + 1.0 if self.d % 2 == 0 else smoothness + 0.5
        alpha = self.nu + self.d / 2.0
        assert abs(alpha - round(alpha)) < 1e-12
        self.alpha = int(round(alpha))
        self.pattern = matern_pattern(self.g, self.alpha)
        self.colptr = self.pattern.indptr.astype(np.int64)
        self.rowval = self.pattern.indices.astype(np.int64)

    def precision(self, tau: float = 1.0, range_: float = 0.3, kappa: float | None = None) -> sp.csc_matrix:
        nu, d = self.nu, self.d
        if kappa is None:
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

    # -- O(nnz) re-evaluation for hyperparameter loops -----------------------------------------------------------
    # K C^-1 K ... K (alpha factors) with K = kappa^2 C + G expands binomially:
    #     Q(tau, kappa) = tau * ratio(kappa) * sum_j binom(alpha, j) kappa^(2 (alpha - j)) B_j,   B_0 = C, B_j = G (C^-1 G)^(j-1)
    # so the nzval of any (tau, range) is a linear combination of alpha + 1 fixed arrays laid out on the structural
    # pattern -- the fixed-pattern value assembly of matern_spde.jl:332-356 / fem_utils.jl:313-335 without sparse
    # products, cheap enough to run on the device right before a refactorization.
    def basis(self) -> np.ndarray:
        if getattr(self, "_basis", None) is None:
            cinv = sp.diags(1.0 / self.c)
            mats = [sp.csc_matrix(sp.diags(self.c)), sp.csc_matrix(self.g)]
            for _ in range(2, self.alpha + 1):
                mats.append(sp.csc_matrix(mats[-1] @ cinv @ self.g))
            self._basis = np.stack([_scatter_into_pattern(self.pattern, m) for m in mats[: self.alpha + 1]])
        return self._basis

    def coefficients(self, tau: float = 1.0, range_: float = 0.3) -> np.ndarray:
        nu, d, a = self.nu, self.d, self.alpha
        kappa = math.sqrt(8.0 * nu) / range_
        ratio = math.gamma(nu) / (math.gamma(nu + d / 2.0) * (4.0 * math.pi) ** (d / 2.0) * kappa ** (2.0 * nu))
        return np.array([tau * ratio * math.comb(a, j) * kappa ** (2 * (a - j)) for j in range(a + 1)])

    def values_from_basis(self, tau: float = 1.0, range_: float = 0.3) -> np.ndarray:
        return self.coefficients(tau, range_) @ self.basis()


def p1_advection(coords: np.ndarray, cells: np.ndarray, gamma) -> sp.csc_matrix:
    """P1 advection matrix for a constant velocity gamma: Be[i, j] = |T| / (d + 1) * gamma . grad(phi_j)
    (assemble_advection_matrix, fem_utils.jl:132-169)."""
    nv, d = coords.shape
    x = coords[cells]
    edges = x[:, 1:, :] - x[:, :1, :]
    vol = np.abs(np.linalg.det(edges)) / math.factorial(d)
    inv = np.linalg.inv(edges)
    grads = np.empty((cells.shape[0], d + 1, d))
    grads[:, 1:, :] = np.transpose(inv, (0, 2, 1))
    grads[:, 0, :] = -grads[:, 1:, :].sum(axis=1)
    gdot = grads @ np.asarray(gamma, dtype=np.float64)              # (ne, d+1): gamma . grad(phi_j)
    be = (vol / (d + 1))[:, None, None] * np.broadcast_to(gdot[:, None, :], (cells.shape[0], d + 1, d + 1))
    rows = np.repeat(cells, d + 1, axis=1).ravel()
    cols = np.tile(cells, (1, d + 1)).ravel()
    b = sp.coo_matrix((be.ravel(), (rows, cols)), shape=(nv, nv)).tocsc()
    b.sum_duplicates()
    b.sort_indices()
    return b


class AdvectionDiffusionSSM:
    """Space-time precision of the implicit-Euler advection-diffusion SPDE (BASELINE config 5), time-major ordering
    (block t = all spatial dofs at time t):
        ext/.../advection_diffusion.jl:103-205   K = kappa^2 M + G (alpha = 1), P = K + B, G_dt = M + dt/c P,
                                                 noise tau/sqrt(c), Q_s / Q_0 = Matern smoothness 1 / 2 with the same H
        implicit_euler_ssm.jl:62-88              Sigma^-1 = M^-1 beta^-1 Q_s beta^-1 M^-1, beta^-1 = (1/sqrt(dt)) / noise
        linear_ssm.jl:63-116                     F^-1 = G' Sigma^-1 G, A'F^-1A = M' Sigma^-1 M, F^-1A = G' Sigma^-1 M;
                                                 diagonal blocks [Q_0 + A'F^-1A ; (F^-1 + A'F^-1A) x (Nt-2) ; F^-1],
                                                 lower off-diagonal blocks -F^-1A
        src/linear_maps/symmetric_block_tridiagonal.jl:77-106  assembled as Symmetric(., :L)
    `posterior(obs_idx, noise_precision)` adds A' Q_eps A for point observations of the first time slice
    (src/arithmetic/condition/linear.jl:53-61)."""

    def __init__(self, coords, cells, nt: int, dt: float = 0.01, kappa: float = 3.0, gamma=(0.3, 0.0), diffusion: float = 0.1,
                 tau: float = 0.1, c: float = 1.0):
        self.ns, self.nt = coords.shape[0], int(nt)
        m, g = p1_mass_stiffness(coords, cells)
        g = sp.csc_matrix(g * diffusion)
        b = p1_advection(coords, cells, gamma)
        M = sp.diags(m)
        P = sp.csc_matrix(kappa ** 2 * M + g + b)
        G = sp.csc_matrix(M + (dt / c) * P)
        q_s = MaternSPDE(coords, cells, 1, diffusion).precision(1.0, kappa=kappa)
        q_0 = MaternSPDE(coords, cells, 2, diffusion).precision(1.0, kappa=kappa)
        beta_inv = math.sqrt(c) / (tau * math.sqrt(dt))
        Minv = sp.diags(1.0 / m)
        sigma_inv = sp.csc_matrix(beta_inv ** 2 * (Minv @ q_s @ Minv))
        Gt_S = sp.csc_matrix(G.T @ sigma_inv)
        F_inv = sp.csc_matrix(Gt_S @ G)
        AtFA = sp.csc_matrix(M @ sigma_inv @ M)
        FA = sp.csc_matrix(Gt_S @ M)
        first = sp.csc_matrix(q_0 + AtFA)
        mid = sp.csc_matrix(F_inv + AtFA)
        blocks = [[None] * self.nt for _ in range(self.nt)]
        for t in range(self.nt):
            blocks[t][t] = first if t == 0 else (mid if t < self.nt - 1 else F_inv)
            if t + 1 < self.nt:
                blocks[t + 1][t] = -FA
                blocks[t][t + 1] = -FA.T
        Q = sp.csc_matrix(sp.bmat(blocks, format="csc"))
        Q = sp.csc_matrix((Q + Q.T) * 0.5)              # Symmetric(., :L): exact symmetry of the assembled values
        Q.sort_indices()
        self.Q = Q
        self.n = Q.shape[0]

    def posterior(self, obs_idx, noise_precision: float) -> sp.csc_matrix:
        d = np.zeros(self.n)
        np.add.at(d, np.asarray(obs_idx, dtype=np.int64), noise_precision)
        Qp = sp.csc_matrix(self.Q + sp.diags(d))
        Qp.sort_indices()
        return Qp


# --------------------------------------------------------------------------- orderings
def geometric_nd_perm(dims, leaf: int = 64, width=1) -> np.ndarray:
    """Geometric nested-dissection permutation of a structured vertex grid `dims`
    (x fastest). `width` = separator thickness (stencil hop count), one number or one per axis (anisotropic stencils:
    a space-time precision couples 5 hops in space but only neighbouring time slices); the cut goes through the axis
    with the most separator-widths left, so cheap (thin-separator) axes are cut first at equal extent. Returns perm with
    perm[k] = original index of the k-th eliminated vertex (0-based). This is a host-side
    ordering a caller may pass as `ordering=perm` (src/workspace/backend.jl:147-153)."""
    dims = tuple(int(v) for v in dims)
    nd = len(dims)
    widths = tuple(int(w) for w in (width if np.ndim(width) else (width,) * nd))
    if len(widths) != nd or min(widths) < 1:
        raise ValueError("width must be a positive number or one positive number per axis")
    strides = np.cumprod((1,) + dims[:-1])
    out = []

    def rec(lo, hi):
        ext = [h - l for l, h in zip(lo, hi)]
        npts = int(np.prod(ext))
        if npts == 0:
            return
        ax = int(np.argmax([e / w for e, w in zip(ext, widths)]))
        width = widths[ax]
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
