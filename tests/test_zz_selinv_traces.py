"""Consumers of the selected inverse: `selinv_dot` with the contraction on the device, traces
against the resident value basis, and the logpdf / logdetcov pullbacks of src/workspace/autodiff.jl:8-91 mirrored in
`gmrf_b200/autodiff.py` -- checked the way test/autodiff/test_logpdf.jl and test_zygote_logdetcov.jl check the
reference's rules: against finite differences and dense `inv` identities.

CPU part (`-m "not gpu"`): the host logic on the dense stand-in backend. GPU part: the same code on the B200 backend,
plus the C-ABI traces against the oracle's selected inverse."""
import numpy as np
import pytest
import scipy.sparse as sp

from dense_backend import DenseBackend
from gmrf_b200 import spde
from gmrf_b200.autodiff import (compute_precision_gradient, logdetcov_basis_gradient, logdetcov_pullback,
                                logpdf_basis_gradient, logpdf_pullback)
from gmrf_b200.workspace import GMRFWorkspace
from gmrf_b200.workspace_gmrf import WorkspaceGMRF


def dense_kw():
    return {"backend_type": DenseBackend}


def gpu_kw():
    from gmrf_b200.backend import B200Backend
    return {"backend_type": B200Backend, "device": 0}


BACKENDS = [pytest.param(dense_kw, id="dense-host-logic"), pytest.param(gpu_kw, id="b200", marks=pytest.mark.gpu)]


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=0.15, random_state=rng, format="csc")
    Q = (A + A.T + sp.identity(n) * (n * 0.5)).tocsc()
    Q.sort_indices()
    return Q


def _sym_perturbation(Q, k, eps):
    """Q + eps * (E_ij + E_ji) at the k-th stored entry (i, j) of the upper triangle, pattern unchanged."""
    C = sp.triu(Q).tocoo()
    i, j = int(C.row[k]), int(C.col[k])
    D = sp.csc_matrix(([eps, eps] if i != j else [eps], ([i, j] if i != j else [i], [j, i] if i != j else [i])), shape=Q.shape)
    Qp = (Q + D).tocsc()
    Qp.sort_indices()
    assert np.array_equal(Qp.indptr, Q.indptr)
    return Qp, i, j


def test_compute_precision_gradient_sparse_matches_dense_formula():
    Q = _spd(12, 0)
    S = sp.csc_matrix(np.linalg.inv(Q.toarray()))
    r = np.random.default_rng(1).standard_normal(12)
    G = compute_precision_gradient(S, r, 0.7)
    assert np.allclose(G.toarray(), 0.5 * 0.7 * (S.toarray() - np.outer(r, r)), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("kw", BACKENDS)
def test_logpdf_pullback_matches_finite_differences(kw):
    n = 14
    Q = _spd(n, 2)
    rng = np.random.default_rng(3)
    mu, z = rng.standard_normal(n), rng.standard_normal(n)
    x = WorkspaceGMRF(mu, Q, **kw())
    ybar = 1.3
    mu_bar, Q_bar, z_bar = logpdf_pullback(x, z, ybar)
    Qd = Q.toarray()
    assert np.allclose(mu_bar, ybar * Qd @ (z - mu), rtol=1e-12)
    assert np.allclose(z_bar, -mu_bar, rtol=1e-12)
    Sigma = np.linalg.inv(Qd)
    Qb = Q_bar.toarray()
    mask = Qd != 0                                         # exact on Q's own pattern (logdetcov.jl:14-18)
    assert np.allclose(Qb[mask], (0.5 * ybar * (Sigma - np.outer(z - mu, z - mu)))[mask], rtol=1e-8, atol=1e-12)
    eps = 1e-6
    for k in (0, 3, 7, 11):
        Qp, i, j = _sym_perturbation(Q, k, eps)
        Qm, _, _ = _sym_perturbation(Q, k, -eps)
        fd = (WorkspaceGMRF(mu, Qp, **kw()).logpdf(z) - WorkspaceGMRF(mu, Qm, **kw()).logpdf(z)) / (2 * eps)
        want = Qb[i, j] + (Qb[j, i] if i != j else 0.0)
        assert abs(ybar * fd - want) <= 1e-6 * max(1.0, abs(want))
    for i in (0, 5):
        e = np.zeros(n)
        e[i] = eps
        fd = (WorkspaceGMRF(mu + e, Q, **kw()).logpdf(z) - WorkspaceGMRF(mu - e, Q, **kw()).logpdf(z)) / (2 * eps)
        assert abs(ybar * fd - mu_bar[i]) <= 1e-6 * max(1.0, abs(mu_bar[i]))


@pytest.mark.parametrize("kw", BACKENDS)
def test_constrained_logpdf_pullback_matches_finite_differences(kw):
    n = 12
    Q = _spd(n, 4)
    rng = np.random.default_rng(5)
    mu = rng.standard_normal(n)
    A = np.vstack([np.ones(n), np.arange(n) % 3 == 0]).astype(float)
    e = np.array([0.3, -0.2])
    x = WorkspaceGMRF(mu, Q, A=A, e=e, **kw())
    z = x.rand(np.random.default_rng(6))                    # a point on the constraint surface
    assert np.allclose(A @ z, e, atol=1e-9)
    mu_bar, Q_bar, _ = logpdf_pullback(x, z, 1.0)
    assert isinstance(Q_bar, np.ndarray) and Q_bar.shape == (n, n)
    eps = 1e-6
    for k in (1, 6, 9):
        Qp, i, j = _sym_perturbation(Q, k, eps)
        Qm, _, _ = _sym_perturbation(Q, k, -eps)
        fd = (WorkspaceGMRF(mu, Qp, A=A, e=e, **kw()).logpdf(z) - WorkspaceGMRF(mu, Qm, A=A, e=e, **kw()).logpdf(z)) / (2 * eps)
        want = Q_bar[i, j] + (Q_bar[j, i] if i != j else 0.0)
        assert abs(fd - want) <= 1e-6 * max(1.0, abs(want))
    for i in (2, 7):
        d = np.zeros(n)
        d[i] = eps
        fd = (WorkspaceGMRF(mu + d, Q, A=A, e=e, **kw()).logpdf(z) - WorkspaceGMRF(mu - d, Q, A=A, e=e, **kw()).logpdf(z)) / (2 * eps)
        assert abs(fd - mu_bar[i]) <= 1e-6 * max(1.0, abs(mu_bar[i]))


@pytest.mark.parametrize("kw", BACKENDS)
def test_logdetcov_pullback_ignores_constraints(kw):
    """The constrained workspace gives the unconstrained -Q^-1 (the regression the reference pins,
    src/workspace/autodiff.jl:59-66)."""
    n = 12
    Q = _spd(n, 7)
    mu = np.zeros(n)
    xu = WorkspaceGMRF(mu, Q, **kw())
    xc = WorkspaceGMRF(mu, Q, A=np.ones((1, n)), e=np.zeros(1), **kw())
    Sigma = np.linalg.inv(Q.toarray())
    mask = Q.toarray() != 0
    for x in (xu, xc):
        Qb = logdetcov_pullback(x, 2.0).toarray()
        assert np.allclose(Qb[mask], (-2.0 * Sigma)[mask], rtol=1e-8)
    assert logdetcov_pullback(xu, 0.0) is None


@pytest.mark.parametrize("kw", BACKENDS)
def test_basis_gradients_of_a_matern_model(kw):
    """Q(tau, range) = sum_j c_j(tau, range) B_j: the contracted pullbacks against tr(Q^-1 B_j) reproduce
    d logdetcov / d tau = -n / tau exactly, d / d range by central differences, and the logpdf gradient."""
    model = spde.MaternSPDE(*spde.mesh2d(10), 1)
    n = model.n
    tau, rho = 0.8, 0.6
    Q = model.precision(tau, rho)
    basis = model.basis()
    c = model.coefficients(tau, rho)
    rng = np.random.default_rng(8)
    mu, z = np.zeros(n), 0.1 * rng.standard_normal(n)
    x = WorkspaceGMRF(mu, Q, **kw())
    x.workspace.backend.set_value_basis(basis)
    g_ld = logdetcov_basis_gradient(x)
    Sigma = np.linalg.inv(Q.toarray())
    cols = np.repeat(np.arange(n), np.diff(Q.indptr))
    tr_dense = basis @ Sigma[Q.indices, cols]
    # tolerance: entries of Sigma to relative 1e-8 (north_star) => a trace to 1e-8 * sum |Sigma_ij B_ij| (the sums cancel)
    tr_abs = np.abs(basis) @ np.abs(Sigma[Q.indices, cols])
    assert np.all(np.abs(g_ld + tr_dense) <= 1e-8 * tr_abs)
    assert abs(c @ g_ld + n) <= 1e-8 * (np.abs(c) @ tr_abs)  # tr(Q^-1 Q) = n  <=>  d logdetcov / d log tau = -n
    h = 1e-5
    dc = (model.coefficients(tau, rho + h) - model.coefficients(tau, rho - h)) / (2 * h)
    ld = lambda r_: -np.linalg.slogdet(model.precision(tau, r_).toarray())[1]
    fd = (ld(rho + h) - ld(rho - h)) / (2 * h)
    assert abs(dc @ g_ld - fd) <= 1e-6 * abs(fd)
    g_lp = logpdf_basis_gradient(x, z, basis, 1.0)
    lp = lambda r_: WorkspaceGMRF(mu, model.precision(tau, r_), **kw()).logpdf(z)
    fd = (lp(rho + h) - lp(rho - h)) / (2 * h)
    assert abs(dc @ g_lp - fd) <= 1e-6 * max(1.0, abs(fd))


# ---------------------------------------------------------------------------------------------- C-ABI traces, GPU only
@pytest.mark.gpu
def test_selinv_dot_device_contraction_vs_oracle():
    import oracle
    Q = spde.MaternSPDE(*spde.mesh3d(7), 0).precision(1.0, 0.5)
    n = Q.shape[0]
    ws = GMRFWorkspace(Q, device=0)
    F = oracle.OracleFactor(Q, ws.backend.permutation())
    Z = F.selinv()                                          # CSC on the (simplicial) factor pattern, original ordering
    rng = np.random.default_rng(9)
    B = Q.copy()
    B.data = rng.standard_normal(B.nnz)                     # not symmetric: dot(Z, B) runs over B as stored
    want = float(Z.multiply(B).sum())
    scale = float(abs(Z.multiply(B)).sum())                 # entries of Sigma to relative 1e-8 => trace to 1e-8 * sum |terms|
    got = ws.selinv_dot(B)
    assert abs(got - want) <= 1e-8 * scale
    host = float(np.dot(ws.selinv_extract_at(B).data, B.data))
    assert abs(got - host) <= 1e-13 * scale                 # same gathered values, different (fixed) summation order
    assert ws.selinv_dot(B) == got                          # bit-reproducible
    assert abs(ws.selinv_dot(Q) - n) <= 1e-8 * float(abs(Z.multiply(Q)).sum())   # tr(Q^-1 Q) = n
    # ragged / degenerate patterns: empty matrix, one entry, entries outside the factor's pattern count 0
    assert ws.selinv_dot(sp.csc_matrix((n, n))) == 0.0
    one = sp.csc_matrix(([2.5], ([3], [3])), shape=(n, n))
    assert abs(ws.selinv_dot(one) - 2.5 * ws.selinv_diag()[3]) <= 1e-14
    R = sp.random(n, n, density=0.01, random_state=rng, format="csc")
    R.sort_indices()
    assert abs(ws.selinv_dot(R) - float(np.dot(ws.selinv_extract_at(R).data, R.data))) <= 1e-12
    with pytest.raises(ValueError):
        ws.selinv_dot(sp.identity(n + 1, format="csc"))


@pytest.mark.gpu
def test_selinv_dot_basis_vs_oracle_and_state_errors():
    import oracle
    from gmrf_b200.backend import B200Backend
    model = spde.MaternSPDE(*spde.mesh2d(40), 1)
    Q = model.precision(1.5, 0.4)
    be = B200Backend(Q, device=0)
    with pytest.raises(RuntimeError):
        be.selinv_dot_basis()                               # no basis uploaded yet
    basis = model.basis()
    be.set_value_basis(basis)
    be.refactorize_combination(model.coefficients(1.5, 0.4))
    tr = be.selinv_dot_basis()
    F = oracle.OracleFactor(Q, be.permutation())
    Z = F.selinv()
    scales = []
    for j in range(basis.shape[0]):
        Bj = sp.csc_matrix((basis[j], Q.indices, Q.indptr), shape=Q.shape)
        want = float(Z.multiply(Bj).sum())
        scale = float(abs(Z.multiply(Bj)).sum())
        scales.append(scale)
        assert abs(tr[j] - want) <= 1e-8 * scale            # Sigma to relative 1e-8 entrywise, the sum cancels
        assert abs(tr[j] - be.selinv_dot(Bj)) <= 1e-13 * scale   # upper triangle twice vs both triangles: order only
    assert abs(model.coefficients(1.5, 0.4) @ tr - model.n) <= 1e-8 * (model.coefficients(1.5, 0.4) @ np.array(scales))
    assert np.array_equal(be.selinv_dot_basis(), tr)        # bit-reproducible
    # a new factorization invalidates Z: the traces follow the new values
    be.refactorize_combination(model.coefficients(0.5, 0.8))
    tr2 = be.selinv_dot_basis()
    assert abs(model.coefficients(0.5, 0.8) @ tr2 - model.n) <= 1e-6 * model.n
    assert not np.array_equal(tr2, tr)
    be.close()


# ------------------------------------------------------------------------------ host replay of the basis-trace gather map
def test_basis_trace_gather_map_on_replayed_panels():
    """`gmrf_b200_selinv_dot_basis` reads Z through the Q -> panel scatter map (the factor and Z share one panel
    layout) with weight 1 on diagonal and 2 on off-diagonal entries. Replayed on the host with the library's own
    tables and a numpy Takahashi recursion: the weighted gather equals tr(Q^-1 B_j)."""
    import replay
    from gmrf_b200 import _lib
    from gmrf_b200.backend import _Handle
    model = spde.MaternSPDE(*spde.mesh2d(12), 1)
    Q = model.precision(1.2, 0.5)
    h = _Handle(Q.shape[0], Q.indptr.astype(np.int64), Q.indices.astype(np.int64), None, _lib.ORDER_ND, device=-1)
    T = replay.Tables(h)
    Zx = replay.selinv(T, replay.factor(T, Q.data))
    diag_nz = np.flatnonzero(Q.indices == np.repeat(np.arange(Q.shape[0]), np.diff(Q.indptr)))
    w = np.where(np.isin(T.q_src, diag_nz), 1.0, 2.0)
    basis = model.basis()
    got = (basis[:, T.q_src] * (w * Zx[T.q_dst])).sum(axis=1)
    Sigma = np.linalg.inv(Q.toarray())
    cols = np.repeat(np.arange(Q.shape[0]), np.diff(Q.indptr))
    want = basis @ Sigma[Q.indices, cols]
    scale = np.abs(basis) @ np.abs(Sigma[Q.indices, cols])
    assert np.all(np.abs(got - want) <= 1e-9 * scale)
    h.close()


# ---------------------------------------------------------------------------------------------- factor export, GPU only
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["matern2d", "matern3d", "tridiag"])
def test_cholesky_sqrt_export(case):
    """R = P'L gathered out of the factor panels on the device: R R' = Q, equal to the oracle's
    `sparse(L)[invperm(p), :]` under the same ordering (sparse_cho_sqrt, src/linear_maps/cholesky_sqrt.jl:6-21), and
    refreshed by a refactorization with the pattern object reused."""
    import oracle
    from gmrf_b200.backend import B200Backend
    Q = {"matern2d": lambda: spde.MaternSPDE(*spde.mesh2d(30), 1).precision(1.0, 0.5),
         "matern3d": lambda: spde.MaternSPDE(*spde.mesh3d(8), 0).precision(1.0, 0.5),
         "tridiag": lambda: spde.tridiag_fixture(10)}[case]()
    Q = sp.csc_matrix(Q)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    R = be.cholesky_sqrt()
    assert abs(R @ R.T - Q).max() <= 1e-12 * abs(Q).max()
    p = be.permutation()
    iperm = np.empty(n, dtype=np.int64)
    iperm[p] = np.arange(n)
    F = oracle.OracleFactor(Q, p)
    Ro = sp.csc_matrix(sp.csc_matrix((F.Lx, F.Li, F.Lp), shape=(n, n))[iperm, :])
    assert abs(R - Ro).max() <= 1e-9 * abs(Ro).max()
    assert abs(2.0 * np.sum(np.log(R[p, np.arange(n)])) - be.compute_logdet()) <= 1e-10 * abs(be.compute_logdet())
    z = np.random.default_rng(0).standard_normal(n)
    assert np.linalg.norm(R.T @ be.backend_backward_solve(z) - z) <= 1e-9 * np.linalg.norm(z)   # L' P x = z
    be.refactorize(2.0 * Q)
    R2 = be.cholesky_sqrt()
    assert R2.indices is R.indices or np.array_equal(R2.indices, R.indices)
    assert abs(R2 - np.sqrt(2.0) * R).max() <= 1e-12 * abs(R).max()
    be.close()


# ---------------------------------------------------------------------------------------------- golden traces, GPU only
@pytest.mark.gpu
@pytest.mark.parametrize("ordering", [None, "amd", "natural"])
def test_selinv_dot_against_golden_traces(ordering):
    """tr(Q^-1 B) of the six golden fixtures (dense LAPACK inverse, tests/golden/make_golden.py) through the device-side
    contraction, under three orderings."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "fixtures.npz"))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name, Q in mg.fixtures().items():
        Qc, cols, bvals = mg.dot_matrix_values(Q)
        B = sp.csc_matrix((bvals, Qc.indices, Qc.indptr), shape=Qc.shape)
        kw = {} if ordering is None else {"ordering": ordering}
        ws = GMRFWorkspace(Q, device=0, **kw)
        got = ws.selinv_dot(B)
        assert abs(got - float(gold[name + "/dot_value"])) <= 1e-8 * float(gold[name + "/dot_scale"]), name
        assert abs(ws.selinv_dot(Q) - Q.shape[0]) <= 1e-8 * float(abs(Q).multiply(abs(ws.selinv_extract_at(Q))).sum()), name


# ---------------------------------------------------------------------------------------------- gradient at the closed-form optimum
@pytest.mark.parametrize("kw", BACKENDS)
def test_basis_gradient_vanishes_at_the_closed_form_scale_estimate(kw):
    """For fixed range, Q = tau * Q_1 and the maximum-likelihood scale is tau_hat = n / (z' Q_1 z). The chain rule through
    the contracted pullback, d logpdf / d log tau = sum_j c_j * d logpdf / d c_j, must vanish there and have the sign of
    (tau_hat - tau) elsewhere."""
    model = spde.MaternSPDE(*spde.mesh2d(9), 1)
    n = model.n
    rho = 0.5
    basis = model.basis()
    Q1 = model.precision(1.0, rho)
    z = np.linalg.solve(np.linalg.cholesky(Q1.toarray() * 2.5).T, np.random.default_rng(3).standard_normal(n))   # ~ N(0, (2.5 Q_1)^-1)
    tau_hat = n / float(z @ (Q1 @ z))
    x = WorkspaceGMRF(np.zeros(n), model.precision(tau_hat, rho), **kw())
    be = x.workspace.backend
    be.set_value_basis(basis)

    def dlogpdf_dlogtau(tau):
        c = model.coefficients(tau, rho)
        d = WorkspaceGMRF(np.zeros(n), model.precision(tau, rho), x.workspace)
        return float(c @ logpdf_basis_gradient(d, z, basis))

    assert abs(dlogpdf_dlogtau(tau_hat)) <= 1e-7 * n          # = 0.5 (n - tau z'Q_1 z): terms of size n/2 cancel
    assert dlogpdf_dlogtau(0.5 * tau_hat) > 0.2 * n and dlogpdf_dlogtau(2.0 * tau_hat) < -0.4 * n
    assert abs(dlogpdf_dlogtau(0.5 * tau_hat) - 0.25 * n) <= 1e-6 * n and abs(dlogpdf_dlogtau(2.0 * tau_hat) + 0.5 * n) <= 1e-6 * n
