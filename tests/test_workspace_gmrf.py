"""WorkspaceGMRF and the workspace Newton loop, mirrored from
test/workspace/test_workspace_gmrf.jl and test/workspace/test_workspace_gaussian_approximation.jl.

CPU part (`-m "not gpu"`): the host logic on a dense stand-in backend against closed forms.
GPU part: the same code on the B200 backend against the dense arm, at the reference's tolerances."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from dense_backend import DenseBackend  # noqa: E402
from gmrf_b200.workspace import GMRFWorkspace  # noqa: E402
from gmrf_b200.workspace_gmrf import ConstraintInfo, PoissonLikelihood, WorkspaceGMRF, gaussian_approximation  # noqa: E402

Y10 = [2, 1, 3, 0, 4, 1, 2, 3, 1, 0]      # test_workspace_gaussian_approximation.jl:17


def tridiag(n, d, e):
    return sp.diags([np.full(n - 1, e), np.full(n, d), np.full(n - 1, e)], [-1, 0, 1]).tocsc()


def grad_norm(Q, mu, lik, x):
    return np.max(np.abs(Q @ (x - mu) - lik.loggrad(x)))


def dense_kw():
    return {"backend_type": DenseBackend}


def gpu_kw():
    from gmrf_b200.backend import B200Backend
    return {"backend_type": B200Backend, "device": 0}


BACKENDS = [pytest.param(dense_kw, id="dense-host-logic"), pytest.param(gpu_kw, id="b200", marks=pytest.mark.gpu)]


# ------------------------------------------------------------------------------------------------ WorkspaceGMRF
@pytest.mark.parametrize("kw", BACKENDS)
def test_workspace_gmrf_matches_dense_identities(kw):
    n = 30
    rng = np.random.default_rng(0)
    Q = (tridiag(n, 2.5, -1.0) + sp.diags(rng.uniform(0.0, 0.5, n))).tocsc()
    mu = rng.standard_normal(n)
    d = WorkspaceGMRF(mu, Q, **kw())
    Qd = Q.toarray()
    Sigma = np.linalg.inv(Qd)
    assert len(d) == n and d.mean() is d.mean_
    assert abs(d.logdetcov() + np.linalg.slogdet(Qd)[1]) <= 1e-10 * abs(d.logdetcov())
    assert np.allclose(d.var(), np.diag(Sigma), rtol=1e-8)
    assert np.allclose(d.std(), np.sqrt(np.diag(Sigma)), rtol=1e-8)
    z = rng.standard_normal(n)
    want = -0.5 * (z - mu) @ Qd @ (z - mu) + 0.5 * np.linalg.slogdet(Qd)[1] - 0.5 * n * np.log(2 * np.pi)
    assert abs(d.logpdf(z) - want) <= 1e-10 * abs(want)
    # a model-structure precomputed log-determinant answers without the workspace (workspace_gmrf.jl:252-258)
    d2 = WorkspaceGMRF(mu, Q, d.workspace, precision_logdet=123.0)
    assert d2.logdetcov() == -123.0


@pytest.mark.parametrize("kw", BACKENDS)
def test_shared_workspace_reloads_the_owner(kw):
    n = 12
    Q1, Q2 = tridiag(n, 2.0, -0.8), tridiag(n, 3.0, -0.5)
    ws = GMRFWorkspace(Q1, **kw())
    a, b = WorkspaceGMRF(np.zeros(n), Q1, ws), WorkspaceGMRF(np.ones(n), Q2, ws)
    for _ in range(2):                                   # alternate owners: ensure_loaded! must reload + refactorize
        assert np.allclose(a.var(), np.diag(np.linalg.inv(Q1.toarray())), rtol=1e-8)
        assert np.allclose(b.var(), np.diag(np.linalg.inv(Q2.toarray())), rtol=1e-8)
        assert abs(a.logdetcov() + np.linalg.slogdet(Q1.toarray())[1]) < 1e-9
    with pytest.raises(ValueError):
        WorkspaceGMRF(np.zeros(n), sp.identity(n, format="csc"), ws)        # pattern mismatch


@pytest.mark.parametrize("kw", BACKENDS)
def test_constrained_workspace_gmrf(kw):
    n = 20
    Q = tridiag(n, 2.2, -1.0)
    mu = np.linspace(-1, 1, n)
    A = np.vstack([np.ones(n), np.r_[np.ones(5), np.zeros(n - 5)]])          # sum-to-zero + a partial sum
    e = np.array([0.0, 0.3])
    ws = GMRFWorkspace(Q, **kw())
    d = WorkspaceGMRF(mu, Q, ws, A, e)
    Sigma = np.linalg.inv(Q.toarray())
    K = Sigma @ A.T @ np.linalg.inv(A @ Sigma @ A.T)
    mean_c = mu - K @ (A @ mu - e)
    Sigma_c = Sigma - K @ A @ Sigma
    assert np.allclose(d.mean(), mean_c, rtol=1e-9, atol=1e-12)
    assert np.allclose(A @ d.mean(), e, atol=1e-10)
    assert np.allclose(d.var(), np.diag(Sigma_c), rtol=1e-7, atol=1e-12)
    rng = np.random.default_rng(3)
    x = d.rand(rng)
    assert np.allclose(A @ x, e, atol=1e-9)                                   # samples satisfy the constraint
    X = d.rand(rng, 4000)
    assert X.shape == (n, 4000) and np.allclose(A @ X, e[:, None], atol=1e-9)
    assert np.allclose(X.var(axis=1), np.diag(Sigma_c), rtol=0.2, atol=2e-3)
    # log-density correction (Rue & Held 2005, 2.3.3; workspace_gmrf.jl:47-51)
    z = mean_c
    base = WorkspaceGMRF(mu, Q, ws).logpdf(z)
    resid = e - A @ mu
    S = A @ Sigma @ A.T
    corr = 0.5 * (2 * np.log(2 * np.pi) + np.linalg.slogdet(S)[1] + resid @ np.linalg.solve(S, resid)) \
        - 0.5 * np.linalg.slogdet(A @ A.T)[1]
    assert abs(d.logpdf(z) - (base + corr)) <= 1e-9 * abs(base + corr)
    with pytest.raises(ValueError):
        ConstraintInfo(ws, mu, np.ones((1, n + 1)), [0.0])


@pytest.mark.parametrize("kw", BACKENDS)
def test_batched_rand_is_the_sequential_stream(kw):
    n = 25
    d = WorkspaceGMRF(np.arange(n, dtype=float), tridiag(n, 2.0, -0.7), **kw())
    seq_rng, blk_rng = np.random.default_rng(42), np.random.default_rng(42)
    seq = np.column_stack([d.rand(seq_rng) for _ in range(11)])              # workspace_gmrf.jl:275-286, one at a time
    blk = d.rand(blk_rng, 11)                                                 # one blocked half solve (wide path on the GPU)
    assert np.allclose(seq, blk, rtol=1e-11, atol=1e-12)
    Sigma = np.linalg.inv(d.precision.toarray())
    X = d.rand(np.random.default_rng(1), 20000)
    assert np.allclose(np.cov(X), Sigma, atol=0.03)                           # samples matched in distribution
    assert np.allclose(X.mean(axis=1), d.mean(), atol=0.03)


@pytest.mark.parametrize("kw", BACKENDS)
def test_linear_predictor_variances_read_sigma_locally(kw):
    """diag(A Sigma A') through selinv_extract_at on the pattern of A'A (linear_predictor_marginals.jl:137-165), with and
    without constraints, against the dense product."""
    from gmrf_b200.linear_predictor import linear_predictor_variances
    n = 40
    rng = np.random.default_rng(8)
    Q = (tridiag(n, 2.4, -1.0) + sp.diags(rng.uniform(0, 0.4, n))).tocsc()
    rows = np.repeat(np.arange(15), 2)
    cols = np.stack([rng.integers(0, n - 1, 15), np.zeros(15, dtype=int)], axis=1)
    cols[:, 1] = cols[:, 0] + 1                                   # each observation touches two NEIGHBOURING sites
    A = sp.csr_matrix((rng.standard_normal(30), (rows, cols.ravel())), shape=(15, n))
    Sigma = np.linalg.inv(Q.toarray())
    ga = WorkspaceGMRF(np.zeros(n), Q, **kw())
    v = linear_predictor_variances(ga, A)
    assert np.allclose(v, np.einsum("ij,jk,ik->i", A.toarray(), Sigma, A.toarray()), rtol=1e-8)
    C = np.ones((1, n))
    gc = WorkspaceGMRF(np.zeros(n), Q, ga.workspace, C, [0.0])
    K = Sigma @ C.T @ np.linalg.inv(C @ Sigma @ C.T)
    Sigma_c = Sigma - K @ C @ Sigma
    vc = linear_predictor_variances(gc, A)
    assert np.allclose(vc, np.einsum("ij,jk,ik->i", A.toarray(), Sigma_c, A.toarray()), rtol=1e-7, atol=1e-12)
    with pytest.raises(ValueError):
        linear_predictor_variances(ga, sp.csr_matrix((3, n + 1)))


# ------------------------------------------------------------------------------------------------ Newton loop
@pytest.mark.parametrize("kw", BACKENDS)
def test_poisson_ga_matches_dense_arm(kw):
    n = 10
    Q = tridiag(n, 2.0, -0.8)
    mu = np.zeros(n)
    lik = PoissonLikelihood(Y10)
    ref = gaussian_approximation(WorkspaceGMRF(mu, Q, **dense_kw()), lik)
    prior = WorkspaceGMRF(mu, Q, **kw())
    res = gaussian_approximation(prior, lik)
    assert isinstance(res, WorkspaceGMRF)
    # the final factorization is deferred (#162): values set, not yet factorized; the first consumer builds it
    assert not res.workspace.numeric_valid
    assert np.allclose(res.mean(), ref.mean(), rtol=1e-8)
    assert np.allclose(res.precision.toarray(), ref.precision.toarray(), rtol=1e-8)
    assert abs(res.logpdf(res.mean()) - ref.logpdf(ref.mean())) <= 1e-6 * abs(ref.logpdf(ref.mean()))
    # the posterior precision is Q_prior + diag(exp(mode)); the mode is a stationary point to the loop's tolerance
    assert np.allclose(res.precision.toarray(), Q.toarray() + np.diag(np.exp(res.mean())), rtol=1e-12)
    assert grad_norm(Q, mu, lik, res.mean()) < 1e-3
    v = res.var()
    assert np.all(v > 0) and res.workspace.numeric_valid and np.isfinite(res.logdetcov())
    assert np.allclose(v, np.diag(np.linalg.inv(res.precision.toarray())), rtol=1e-8)
    # the prior's own snapshot is untouched
    assert np.array_equal(prior.precision.toarray(), Q.toarray())


@pytest.mark.parametrize("kw", BACKENDS)
def test_device_side_iterates_follow_the_host_path(kw):
    n = 60
    Q = tridiag(n, 2.01, -1.0)
    lik = PoissonLikelihood(np.round(np.exp(2.0 * np.sin(np.linspace(0.0, 4.0 * np.pi, n)))))
    host, dev = {}, {}
    a = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik, stats=host)
    b = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik, stats=dev, device_iterates=True)
    assert host == dev
    assert np.allclose(a.mean(), b.mean(), rtol=1e-13, atol=1e-15)
    assert np.array_equal(a.precision.data, b.precision.data)
    assert np.allclose(b.var(), np.diag(np.linalg.inv(b.precision.toarray())), rtol=1e-8)


@pytest.mark.parametrize("kw", BACKENDS)
def test_repeated_ga_reuses_the_workspace(kw):
    n = 10
    Q = tridiag(n, 2.0, -0.8)
    lik = PoissonLikelihood(Y10)
    ws = GMRFWorkspace(Q, **kw())
    for scale in (1.0, 2.0, 0.5):                       # outer hyperparameter loop: same pattern, new values
        Qs = (Q * scale).tocsc()
        ws.update_precision(Qs)
        post = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Qs, ws), lik)
        ref = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Qs, **dense_kw()), lik)
        assert np.allclose(post.mean(), ref.mean(), rtol=1e-6)
        assert np.allclose(post.precision.toarray(), ref.precision.toarray(), rtol=1e-6)


@pytest.mark.parametrize("kw", BACKENDS)
def test_partially_observed_field(kw):
    n = 10
    Q = tridiag(n, 2.0, -0.8)
    lik = PoissonLikelihood([2, 3, 4, 1, 2], indices=[0, 2, 4, 6, 8])
    res = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik)
    ref = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **dense_kw()), lik)
    assert np.allclose(res.mean(), ref.mean(), rtol=1e-6)
    assert np.allclose(res.precision.toarray(), ref.precision.toarray(), rtol=1e-6)
    unobserved = [1, 3, 5, 7, 9]
    assert np.array_equal(res.precision.diagonal()[unobserved], Q.diagonal()[unobserved])


@pytest.mark.parametrize("kw", BACKENDS)
def test_step_recovery_policies_and_numerical_floor(kw):
    n = 200                                             # test_workspace_gaussian_approximation.jl:176-217
    Q = tridiag(n, 2.01, -1.0)
    mu = np.zeros(n)
    counts = np.round(np.exp(3.0 * np.sin(np.linspace(0.0, 6.0 * np.pi, n))))
    lik = PoissonLikelihood(counts)
    tight = dict(newton_dec_tol=1e-12, mean_change_tol=1e-12)
    for policy in ("retry_full", "sqrt"):
        ref = gaussian_approximation(WorkspaceGMRF(mu, Q, **dense_kw()), lik, step_recovery=policy, **tight)
        res = gaussian_approximation(WorkspaceGMRF(mu, Q, **kw()), lik, step_recovery=policy, **tight)
        assert np.allclose(res.mean(), ref.mean(), rtol=1e-8)
    ws_sqrt = gaussian_approximation(WorkspaceGMRF(mu, Q, **kw()), lik, max_iter=8, step_recovery="sqrt", **tight)
    ws_full = gaussian_approximation(WorkspaceGMRF(mu, Q, **kw()), lik, max_iter=8, **tight)
    assert grad_norm(Q, mu, lik, ws_full.mean()) < 1e-3 * grad_norm(Q, mu, lik, ws_sqrt.mean())
    conv = gaussian_approximation(WorkspaceGMRF(mu, Q, **kw()), lik, **tight)
    assert grad_norm(Q, mu, lik, conv.mean()) < 1e-12                         # the numerical floor at the mode
    with pytest.raises(ValueError):
        gaussian_approximation(WorkspaceGMRF(mu, Q, **kw()), lik, step_recovery="bogus")


@pytest.mark.parametrize("kw", BACKENDS)
def test_warm_start_and_refactorization_count(kw):
    n = 200                                             # test_predictive_convergence.jl:60-68 (warm-start variant)
    Q = (1.1 * tridiag(n, 2.01, -1.0)).tocsc()
    counts = np.round(np.exp(2.0 * np.sin(np.linspace(0.0, 4.0 * np.pi, n))))
    lik = PoissonLikelihood(counts)
    cold, warm = {}, {}
    first = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik, stats=cold)
    again = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik, x0=first.mean(), stats=warm)
    assert warm["refactorizations"] < cold["refactorizations"]               # a converged start needs fewer factorizations
    assert np.allclose(again.mean(), first.mean(), rtol=1e-4, atol=1e-6)
    # reruns are bit-reproducible (test_predictive_convergence.jl:20-28)
    rerun = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik)
    assert np.array_equal(rerun.mean(), first.mean())


@pytest.mark.gpu
def test_newton_loop_on_a_2d_matern_field():
    """BASELINE config 2 in miniature: Poisson observations on a 2D Matern (alpha = 3) field, repeated numeric
    refactorization on a fixed pattern; the B200 loop must follow the dense arm and end at a stationary point."""
    from gmrf_b200 import spde
    from gmrf_b200.backend import B200Backend
    coords, cells = spde.mesh2d(20)
    Q = spde.MaternSPDE(coords, cells, smoothness=1).precision(1.0, 0.6)
    n = Q.shape[0]
    rng = np.random.default_rng(1)
    lam = np.exp(0.5 + 0.5 * np.sin(2 * np.pi * coords[:, 0]) * np.cos(2 * np.pi * coords[:, 1]))
    lik = PoissonLikelihood(rng.poisson(lam))
    stats = {}
    res = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, backend_type=B200Backend, device=0), lik,
                                 newton_dec_tol=1e-10, mean_change_tol=1e-10, stats=stats)
    ref = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **dense_kw()), lik, newton_dec_tol=1e-10, mean_change_tol=1e-10)
    assert np.allclose(res.mean(), ref.mean(), rtol=1e-7, atol=1e-9)
    assert grad_norm(Q, np.zeros(n), lik, res.mean()) < 1e-8 * max(1.0, np.abs(Q).max())
    assert 2 <= stats["refactorizations"] <= 20
    assert np.allclose(res.std(), ref.std(), rtol=1e-7)
