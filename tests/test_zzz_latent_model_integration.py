"""`(model)(ws; theta...)`, joint-pattern workspaces, `_pad_to_workspace_pattern` and the sparse-Hessian Newton path
(SURVEY.md 8a rows a11, a12), mirrored from test/workspace/test_workspace_latent_models.jl and
test/workspace/test_precision_logdet.jl (hook values, workspace priors carry the hook).

CPU part (`-m "not gpu"`): the host logic on the dense stand-in backend against closed forms.
GPU part: the same code on the B200 backend. (The file sorts last on purpose: its GPU arms were written after round 1's
GPU budget was spent and have only run on the CPU arm so far.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from dense_backend import DenseBackend
from latent_stand_ins import AR1Model, IIDModel, LinearGaussianLikelihood, MaternModel, RW1Model
from gmrf_b200 import spde
from gmrf_b200.latent_model_integration import (copy_values_into, evaluate_with_workspace, make_workspace,
                                                make_workspace_pool, ones_pattern, pad_to_workspace_pattern, workspace_for)
from gmrf_b200.workspace_gmrf import PoissonLikelihood, WorkspaceGMRF, gaussian_approximation


def dense_kw():
    return {"backend_type": DenseBackend}


def gpu_kw():
    from gmrf_b200.backend import B200Backend
    return {"backend_type": B200Backend, "device": 0}


BACKENDS = [pytest.param(dense_kw, id="dense-host-logic"), pytest.param(gpu_kw, id="b200", marks=pytest.mark.gpu)]


def _dense_logpdf(Q, mu, z):
    Qd = Q.toarray()
    r = z - mu
    return -0.5 * r @ Qd @ r + 0.5 * np.linalg.slogdet(Qd)[1] - 0.5 * len(z) * np.log(2 * np.pi)


@pytest.mark.parametrize("kw", BACKENDS)
def test_workspaces_from_models(kw):                     # test_workspace_latent_models.jl:10-29
    assert make_workspace(AR1Model(20), kw(), tau=1.0, rho=0.5).dimension() == 20
    assert make_workspace(RW1Model(15), kw(), tau=1.0).dimension() == 15
    assert make_workspace(IIDModel(10), kw(), tau=2.0).dimension() == 10


@pytest.mark.parametrize("kw", BACKENDS)
def test_model_on_workspace_matches_fresh_construction(kw):     # :31-67
    n = 20
    model = AR1Model(n)
    ws = make_workspace(model, kw(), tau=1.0, rho=0.5)
    z = np.random.default_rng(0).standard_normal(n)
    for tau, rho in ((2.0, 0.3), (1.0, 0.3), (2.0, 0.5), (0.5, 0.8)):
        d = evaluate_with_workspace(model, ws, tau=tau, rho=rho)
        assert isinstance(d, WorkspaceGMRF) and len(d) == n and d.workspace is ws
        Q = model.precision_matrix(tau, rho)
        assert abs(d.logpdf(z) - _dense_logpdf(Q, np.zeros(n), z)) <= 1e-8 * abs(d.logpdf(z))
        assert np.allclose(d.mean(), 0.0)
        assert np.allclose(d.var(), np.diag(np.linalg.inv(Q.toarray())), rtol=1e-8)


@pytest.mark.parametrize("kw", BACKENDS)
def test_constrained_model_on_workspace(kw):             # :69-91
    n = 15
    model = RW1Model(n)
    ws = make_workspace(model, kw(), tau=1.0)
    d = evaluate_with_workspace(model, ws, tau=2.0)
    assert d.has_constraints() and len(d) == n
    assert abs(np.sum(d.mean())) <= 1e-8
    Sigma = np.linalg.inv(model.precision_matrix(2.0).toarray())
    A = np.ones((1, n))
    Sc = Sigma - Sigma @ A.T @ np.linalg.solve(A @ Sigma @ A.T, A @ Sigma)      # constrained covariance
    # base variance (~1 / (n * regularization) = 6.7e3) minus the correction cancels to O(1), and cond(Q) ~ 1e6: the
    # north-star tolerance of 1e-8 applies to the base variances, i.e. absolutely on that scale
    assert np.allclose(d.var(), np.diag(Sc), rtol=1e-8, atol=1e-8 * np.max(np.diag(Sigma)))
    x = d.rand(np.random.default_rng(1))
    assert abs(np.sum(x)) <= 1e-7


@pytest.mark.parametrize("kw", BACKENDS)
def test_inla_like_pipeline(kw):                         # :93-114
    n = 20
    model = AR1Model(n)
    ws = make_workspace(model, kw(), tau=1.0, rho=0.5)
    lik = PoissonLikelihood(np.random.default_rng(2).integers(1, 6, n))
    for tau, rho in ((1.0, 0.3), (2.0, 0.5), (5.0, 0.8)):
        prior = evaluate_with_workspace(model, ws, tau=tau, rho=rho)
        post = gaussian_approximation(prior, lik)
        assert np.isfinite(post.logpdf(post.mean()))
        ref = gaussian_approximation(WorkspaceGMRF(np.zeros(n), model.precision_matrix(tau, rho), **dense_kw()), lik)
        assert np.allclose(post.mean(), ref.mean(), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("kw", BACKENDS)
def test_joint_pattern_workspace_diagonal_hessian(kw):   # :116-130
    n = 20
    model = AR1Model(n)
    lik = PoissonLikelihood(np.random.default_rng(3).integers(1, 6, n))
    ws = workspace_for(model, lik, kw(), tau=1.0, rho=0.5)
    assert ws.dimension() == n and ws.Q.nnz == model.precision_matrix(1.0, 0.5).nnz      # diagonal adds nothing
    post = gaussian_approximation(evaluate_with_workspace(model, ws, tau=2.0, rho=0.3), lik)
    assert np.isfinite(post.logpdf(post.mean()))


@pytest.mark.parametrize("kw", BACKENDS)
def test_joint_pattern_workspace_non_diagonal_hessian(kw):      # :132-165
    n, m_obs = 8, 10
    rng = np.random.default_rng(123)
    A = sp.random(m_obs, n, density=0.6, random_state=rng, format="csr")
    y = rng.standard_normal(m_obs)
    model = AR1Model(n)
    lik = LinearGaussianLikelihood(A, y, 0.5)
    ws = workspace_for(model, lik, kw(), tau=1.0, rho=0.5)
    Q_prior_only = model.precision_matrix(1.0, 0.5)
    assert ws.Q.nnz > Q_prior_only.nnz                           # strictly larger pattern
    assert np.allclose(ws.Q.toarray(), Q_prior_only.toarray())   # ... holding the prior's values
    prior = evaluate_with_workspace(model, ws, tau=2.0, rho=0.3)     # padded, not rejected by the pattern check
    Q = model.precision_matrix(2.0, 0.3)
    z = rng.standard_normal(n)
    assert abs(prior.logpdf(z) - _dense_logpdf(Q, np.zeros(n), z)) <= 1e-8 * abs(prior.logpdf(z))
    post = gaussian_approximation(prior, lik)
    assert np.isfinite(post.logpdf(post.mean()))
    # Gaussian likelihood: the Gaussian approximation is the exact posterior
    Qp = Q.toarray() + (A.T @ A).toarray() / 0.25
    assert np.allclose(post.mean(), np.linalg.solve(Qp, A.T @ y / 0.25), rtol=1e-8, atol=1e-10)
    assert np.allclose(post.precision.toarray(), Qp, rtol=1e-12, atol=1e-12)
    assert np.allclose(post.var(), np.diag(np.linalg.inv(Qp)), rtol=1e-8)


def test_pad_and_copy_helpers():
    n = 6
    model = AR1Model(n)
    ws = make_workspace(model, dense_kw(), tau=1.0, rho=0.5)
    Q = model.precision_matrix(2.0, 0.1)
    assert pad_to_workspace_pattern(Q, ws) is not None and pad_to_workspace_pattern(Q, ws).nnz == ws.Q.nnz
    D = sp.csc_matrix(sp.diags(np.arange(1.0, n + 1)))           # sub-pattern: padded with explicit zeros
    P = pad_to_workspace_pattern(D, ws)
    assert np.array_equal(P.indptr, ws.Q.indptr) and np.array_equal(P.indices, ws.Q.indices)
    assert np.array_equal(P.toarray(), D.toarray()) and P.nnz == ws.Q.nnz
    outside = D.tolil()
    outside[0, n - 1] = 1.0
    with pytest.raises(ValueError, match=r"nonzero at \(1, 6\) outside the workspace pattern"):
        pad_to_workspace_pattern(outside.tocsc(), ws)
    with pytest.raises(ValueError, match="workspace expects"):
        pad_to_workspace_pattern(sp.identity(n + 1, format="csc"), ws)
    J = ones_pattern(ws.Q)
    assert np.all(J.data == 1.0) and np.array_equal(J.indices, ws.Q.indices)
    copy_values_into(J, D)
    assert np.array_equal(J.toarray(), D.toarray()) and J.nnz == ws.Q.nnz


def test_precision_logdet_hook_skips_the_factorization():     # test_precision_logdet.jl:28-89
    n = 12
    model = IIDModel(n)
    ws = make_workspace(model, dense_kw(), tau=1.0)
    before = ws.backend.refactorizations
    d = evaluate_with_workspace(model, ws, tau=3.0)
    assert d.precision_logdet == pytest.approx(n * np.log(3.0))
    assert d.logdetcov() == pytest.approx(-n * np.log(3.0), rel=1e-14)
    assert ws.backend.refactorizations == before                 # answered by the hook: nothing factorized
    z = np.random.default_rng(4).standard_normal(n)
    assert d.logpdf(z) == pytest.approx(_dense_logpdf(model.precision_matrix(3.0), np.zeros(n), z), rel=1e-12)
    assert ws.backend.refactorizations == before
    ar = AR1Model(n)
    ws2 = make_workspace(ar, dense_kw(), tau=1.0, rho=0.2)
    d2 = evaluate_with_workspace(ar, ws2, tau=2.0, rho=0.6)
    assert d2.logdetcov() == pytest.approx(-np.linalg.slogdet(ar.precision_matrix(2.0, 0.6).toarray())[1], rel=1e-12)


@pytest.mark.parametrize("kw", BACKENDS)
def test_values_assembled_on_the_device_give_the_same_object(kw):
    model = MaternModel(spde.MaternSPDE(*spde.mesh2d(8), 1))
    ws = make_workspace(model, kw(), tau=1.0, range_=0.5)
    z = 0.1 * np.random.default_rng(5).standard_normal(model.n)
    for tau, rng_ in ((0.7, 0.4), (2.0, 0.9)):
        host = evaluate_with_workspace(model, ws, tau=tau, range_=rng_)
        lp_host, var_host = host.logpdf(z), host.var().copy()
        dev = evaluate_with_workspace(model, ws, on_device=True, tau=tau, range_=rng_)
        assert ws.numeric_valid and ws.loaded_version == dev.version      # factorized already, owned by `dev`
        assert np.allclose(dev.precision.data, host.precision.data, rtol=1e-13)
        assert abs(dev.logpdf(z) - lp_host) <= 1e-10 * abs(lp_host)
        assert np.allclose(dev.var(), var_host, rtol=1e-8)
    if kw is dense_kw:
        before = ws.backend.refactorizations
        d = evaluate_with_workspace(model, ws, on_device=True, tau=1.1, range_=0.6)
        d.logpdf(z)
        d.var()
        assert ws.backend.refactorizations == before + 1             # one factorization per theta, no reload
    with pytest.raises(ValueError):
        evaluate_with_workspace(model, make_workspace(AR1Model(model.n), kw(), tau=1.0, rho=0.1), on_device=True, tau=1.0, range_=0.5)


def test_pool_from_model():                               # test_workspace_pool.jl:117-125
    model = AR1Model(16)
    pool = make_workspace_pool(model, size=2, backend_kwargs=dense_kw(), tau=1.0, rho=0.5)
    assert len(pool.workspaces) == 2
    z = np.random.default_rng(6).standard_normal(16)
    with pool.with_workspace() as ws:
        d = evaluate_with_workspace(model, ws, tau=2.0, rho=0.4)
        assert abs(d.logpdf(z) - _dense_logpdf(model.precision_matrix(2.0, 0.4), np.zeros(16), z)) <= 1e-10 * abs(d.logpdf(z))
    assert pool.checkout() is not None


# ---------------------------------------------------------------------------------------------- test_workspace_pool.jl
@pytest.mark.parametrize("kw", BACKENDS)
def test_workspace_pool_protocol(kw):
    import threading
    from gmrf_b200.workspace import GMRFWorkspace, WorkspacePool
    n = 30
    rng = np.random.default_rng(7)
    A = sp.random(n, n, density=0.1, random_state=rng, format="csc")
    Q = sp.csc_matrix(A + A.T + 10.0 * sp.identity(n))
    Q.sort_indices()
    Qd = Q.toarray()
    base = kw()
    pkw = lambda: {**{k: v for k, v in base.items() if k != "device"}, "devices": (base.get("device", 0),)}
    pool = WorkspacePool(Q, size=2, **pkw())                     # :15-33 checkout / checkin
    assert len(pool.workspaces) == 2
    ws1, ws2 = pool.checkout(), pool.checkout()
    assert isinstance(ws1, GMRFWorkspace) and ws1.dimension() == n and ws1 is not ws2
    pool.checkin(ws1)
    pool.checkin(ws2)
    ws3 = pool.checkout()
    pool.checkin(ws3)
    with pool.with_workspace() as ws:                            # :35-45 RAII
        assert np.isfinite(np.linalg.norm(ws.workspace_solve(rng.standard_normal(n))))
    one = WorkspacePool(Q, size=1, **pkw())                       # :47-62 returned on exception
    with pytest.raises(RuntimeError):
        with one.with_workspace():
            raise RuntimeError("intentional error")
    one.checkin(one.checkout())
    b = rng.standard_normal(n)                                   # :64-86 independent workspaces
    three = WorkspacePool(Q, size=3, **pkw())
    for i in (1, 2, 3):
        ws = three.checkout()
        ws.update_precision(Q * float(i))
        x = ws.workspace_solve(b)
        three.checkin(ws)
        assert np.allclose(x, np.linalg.solve(Qd * i, b), rtol=1e-10, atol=1e-13)
    results, errs = [None] * 20, []                             # :88-115 parallel correctness

    def work(i):
        try:
            with three.with_workspace() as ws:
                ws.update_precision(Q * float(i + 1))
                results[i] = np.linalg.norm(ws.workspace_solve(np.ones(n)))
        except Exception as e:       # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(20)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for i in range(20):
        assert abs(results[i] - np.linalg.norm(np.linalg.solve(Qd * (i + 1), np.ones(n)))) <= 1e-10 * results[i]


# ------------------------------------------------------------------- plain GMRF prior: the cache-backed Newton loop (a13)
def _gmrf_kw(kw):
    base = kw()
    return {"backend_type": base["backend_type"]} if base["backend_type"] is DenseBackend else {"device": base["device"]}


@pytest.mark.parametrize("kw", BACKENDS)
def test_gaussian_approximation_of_a_plain_gmrf(kw):
    """test/workspace/test_workspace_gaussian_approximation.jl:8-32 run the other way round: the plain `GMRF` path
    (condition/gaussian_approximation.jl:197-229) against the workspace path and against a dense Newton iteration."""
    from gmrf_b200.gmrf import GMRF, gaussian_approximation as ga_gmrf
    n = 10
    Q = sp.diags([np.full(n - 1, -0.8), np.full(n, 2.0), np.full(n - 1, -0.8)], [-1, 0, 1]).tocsc()
    y = np.array([2, 1, 3, 0, 4, 1, 2, 3, 1, 0], dtype=float)
    lik = PoissonLikelihood(y)
    prior = GMRF(np.zeros(n), Q, **_gmrf_kw(kw))
    post = ga_gmrf(prior, lik)
    assert isinstance(post, GMRF) and post.linsolve_cache is not prior.linsolve_cache
    ref = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **dense_kw()), lik)
    assert np.allclose(post.mean(), ref.mean(), rtol=1e-8, atol=1e-10)
    assert np.allclose(post.precision.toarray(), ref.precision.toarray(), rtol=1e-8)
    x = np.zeros(n)                                                   # dense Newton to the mode
    Qd = Q.toarray()
    for _ in range(50):
        g = Qd @ x - (y - np.exp(x))
        x = x - np.linalg.solve(Qd + np.diag(np.exp(x)), g)
    assert np.allclose(post.mean(), x, atol=1e-4)
    # the posterior GMRF owns a live factorization of Q_post: var / logdetcov / rand come straight from it
    Qp = post.precision.toarray()
    assert np.allclose(post.var(), np.diag(np.linalg.inv(Qp)), rtol=1e-8)
    assert abs(post.logdetcov() + np.linalg.slogdet(Qp)[1]) <= 1e-10 * abs(post.logdetcov())
    assert post.rand(np.random.default_rng(0), 4).shape == (n, 4)
    # the prior is untouched
    assert np.allclose(prior.var(), np.diag(np.linalg.inv(Qd)), rtol=1e-8)


@pytest.mark.parametrize("kw", BACKENDS)
def test_plain_gmrf_with_a_non_diagonal_hessian(kw):
    """`_ga_resolve_cache` (:98-110): the posterior precision's storage differs from the prior's, the solver is built for
    the joint pattern. Gaussian likelihood => exact posterior, equal to `linear_condition`."""
    from gmrf_b200.gmrf import GMRF, gaussian_approximation as ga_gmrf, linear_condition
    n, m_obs = 9, 7
    rng = np.random.default_rng(11)
    A = sp.random(m_obs, n, density=0.5, random_state=rng, format="csr")
    yv = rng.standard_normal(m_obs)
    Q = AR1Model(n).precision_matrix(1.5, 0.4)
    mu = rng.standard_normal(n)
    prior = GMRF(mu, Q, **_gmrf_kw(kw))
    post = ga_gmrf(prior, LinearGaussianLikelihood(A, yv, 0.5))
    exact = linear_condition(prior, A, 4.0, yv, **_gmrf_kw(kw))
    assert np.allclose(post.mean(), exact.mean(), rtol=1e-8, atol=1e-10)
    assert np.allclose(post.precision.toarray(), exact.precision.toarray(), rtol=1e-12, atol=1e-12)
    assert np.allclose(post.var(), exact.var(), rtol=1e-8)


# --------------------------------------------------------- constrained Gaussian approximation (test_workspace_constrained.jl:88-133)
def _kkt_mode(Q, A, e, grad, hess_diag, x0, iters=100):
    """Dense equality-constrained Newton: minimise 0.5 x'Qx - loglik(x) subject to A x = e."""
    x = x0.copy()
    m = A.shape[0]
    for _ in range(iters):
        H = Q - np.diag(hess_diag(x))
        g = Q @ x - grad(x)
        K = np.block([[H, A.T], [A, np.zeros((m, m))]])
        step = np.linalg.solve(K, np.concatenate([-g, e - A @ x]))[: x.size]
        x = x + step
        if np.linalg.norm(step) < 1e-13:
            break
    return x


@pytest.mark.parametrize("kw", BACKENDS)
@pytest.mark.parametrize("family", ["poisson", "bernoulli"])
def test_constrained_gaussian_approximation(kw, family):
    """The Newton step is projected onto the constraint's tangent space with one blocked multi-RHS solve
    (`_workspace_constrain_step`, workspace/gaussian_approximation.jl:137-149); the mode is the KKT point."""
    from latent_stand_ins import BernoulliLikelihood
    if family == "poisson":
        y = np.array([2, 1, 3, 0, 4, 1, 2, 3], dtype=float)
        lik, max_iter = PoissonLikelihood(y), 50
    else:
        y = np.array([1, 0, 1, 0], dtype=float)
        lik, max_iter = BernoulliLikelihood(y), 20
    n = y.size
    Q = sp.identity(n, format="csc")
    A, e = np.ones((1, n)), np.zeros(1)
    prior = WorkspaceGMRF(np.zeros(n), Q, A=A, e=e, **kw())
    post = gaussian_approximation(prior, lik, max_iter=max_iter, mean_change_tol=1e-10, newton_dec_tol=1e-12)
    assert isinstance(post, WorkspaceGMRF) and post.has_constraints()
    assert abs(np.sum(post.mean())) <= 1e-6
    mode = _kkt_mode(Q.toarray(), A, e, lik.loggrad, lik.loghessian, np.zeros(n))
    assert np.allclose(post.mean(), mode, rtol=1e-6, atol=1e-8)
    # default tolerances, as in the reference's test: same mode to its rtol
    post_d = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, A=A, e=e, **kw()), lik, max_iter=max_iter)
    assert np.allclose(post_d.mean(), mode, rtol=1e-4, atol=1e-5) and abs(np.sum(post_d.mean())) <= 1e-6
    # two constraints with a non-zero right-hand side
    if family == "poisson":
        A2 = np.vstack([np.ones(n), (np.arange(n) % 2 == 0).astype(float)])
        e2 = np.array([0.5, -0.25])
        post2 = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, A=A2, e=e2, **kw()), lik, mean_change_tol=1e-10, newton_dec_tol=1e-12)
        assert np.allclose(A2 @ post2.mean(), e2, atol=1e-8)
        x0 = np.linalg.lstsq(A2, e2, rcond=None)[0]
        assert np.allclose(post2.mean(), _kkt_mode(Q.toarray(), A2, e2, lik.loggrad, lik.loghessian, x0), rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------------------------- analysis export / import on the device
@pytest.mark.gpu
def test_backend_from_exported_analysis_is_bit_identical():
    """`B200Backend(Q, analysis=blob)` skips ordering / etree / supernodes / schedules and must be indistinguishable from
    the backend the blob came from: same permutation, bit-identical log-determinant, solve, half solve and variances."""
    from gmrf_b200.backend import B200Backend
    model = spde.MaternSPDE(*spde.mesh3d(8), 0)
    Q = model.precision(1.0, 0.5)
    n = Q.shape[0]
    a = B200Backend(Q, device=0)
    blob = a.export_analysis()
    b = B200Backend(Q, device=0, analysis=blob)
    assert np.array_equal(a.permutation(), b.permutation()) and a.info()["nnz_l"] == b.info()["nnz_l"]
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal(n)
    assert a.compute_logdet() == b.compute_logdet()
    assert np.array_equal(a.backend_solve(rhs), b.backend_solve(rhs))
    assert np.array_equal(a.backend_backward_solve(rhs), b.backend_backward_solve(rhs))
    assert np.array_equal(a.get_selinv_diag(), b.get_selinv_diag())
    Q2 = model.precision(0.4, 0.9)
    a.refactorize(Q2)
    b.refactorize(Q2)
    assert a.compute_logdet() == b.compute_logdet()
    assert abs(b.compute_logdet() - np.linalg.slogdet(Q2.toarray())[1]) <= 1e-10 * abs(b.compute_logdet())
    with pytest.raises(ValueError):
        B200Backend(spde.MaternSPDE(*spde.mesh3d(7), 0).precision(1.0, 0.5), device=0, analysis=blob)
    a.close()
    b.close()


# ------------------------------------------------------- remaining testsets of test_workspace_gmrf.jl / test_workspace_gaussian_approximation.jl
def _tri(n, d, e):
    return sp.diags([np.full(n - 1, e), np.full(n, d), np.full(n - 1, e)], [-1, 0, 1]).tocsc()


@pytest.mark.parametrize("kw", BACKENDS)
def test_sqmahal_gradlogpdf_and_shared_means(kw):           # test_workspace_gmrf.jl:76-88, :147-160
    from gmrf_b200.workspace import GMRFWorkspace
    n = 12
    rng = np.random.default_rng(0)
    Q = (_tri(n, 2.5, -0.9) + sp.diags(rng.uniform(0, 0.3, n))).tocsc()
    Qd = Q.toarray()
    mu = rng.standard_normal(n)
    wg = WorkspaceGMRF(mu, Q, **kw())
    x = rng.standard_normal(n)
    assert abs(wg.sqmahal(x) - (x - mu) @ Qd @ (x - mu)) <= 1e-10 * wg.sqmahal(x) and abs(wg.sqmahal(mu)) <= 1e-12
    assert np.allclose(wg.gradlogpdf(mu), 0.0, atol=1e-10)
    assert np.allclose(wg.gradlogpdf(x), -Qd @ (x - mu), rtol=1e-10)
    ws = GMRFWorkspace(Q, **kw())
    wa, wb = WorkspaceGMRF(np.zeros(n), Q, ws), WorkspaceGMRF(np.ones(n), Q, ws)
    z = rng.standard_normal(n)
    assert np.array_equal(wa.mean(), np.zeros(n)) and np.array_equal(wb.mean(), np.ones(n))
    for w, m in ((wa, np.zeros(n)), (wb, np.ones(n))):
        assert abs(w.logpdf(z) - _dense_logpdf(Q, m, z)) <= 1e-10 * abs(w.logpdf(z))


@pytest.mark.parametrize("kw", BACKENDS)
def test_prior_and_posterior_stay_coherent_around_a_gaussian_approximation(kw):
    """test_workspace_gmrf.jl:162-204 and test_workspace_gaussian_approximation.jl:76-122: GA leaves the shared workspace
    at Q_post; the prior keeps evaluating against Q_prior, GA seeds from its owner's snapshot, the posterior's
    factorization is deferred to its first consumer, the prior's precision object is untouched."""
    from gmrf_b200.workspace import GMRFWorkspace
    y5 = np.array([2, 1, 3, 0, 4], dtype=float)
    lik5 = PoissonLikelihood(y5)
    Qp = _tri(5, 2.0, -0.3)
    z = np.random.default_rng(1).standard_normal(5)
    z -= z.mean()
    ws = GMRFWorkspace(Qp.copy(), **kw())
    prior = WorkspaceGMRF(np.zeros(5), Qp.copy(), ws)
    lp_ref = _dense_logpdf(Qp, np.zeros(5), z)
    assert abs(prior.logpdf(z) - lp_ref) <= 1e-10 * abs(lp_ref)
    gaussian_approximation(prior, lik5)
    assert abs(prior.logpdf(z) - lp_ref) <= 1e-10 * abs(lp_ref)          # both the quadratic form and the logdet
    # two priors on one workspace: GA on A after B touched the workspace last
    Qa, Qb = _tri(5, 2.0, -0.3), _tri(5, 5.0, -0.7)
    ws2 = GMRFWorkspace(Qa.copy(), **kw())
    wa, wb = WorkspaceGMRF(np.zeros(5), Qa.copy(), ws2), WorkspaceGMRF(np.zeros(5), Qb.copy(), ws2)
    wb.logpdf(np.random.default_rng(2).standard_normal(5))
    post = gaussian_approximation(wa, lik5)
    ref = gaussian_approximation(WorkspaceGMRF(np.zeros(5), Qa.copy(), **dense_kw()), lik5)
    assert np.allclose(post.mean(), ref.mean(), rtol=1e-8, atol=1e-12)
    assert abs(post.logpdf(np.zeros(5)) - ref.logpdf(np.zeros(5))) <= 1e-8 * abs(ref.logpdf(np.zeros(5)))
    # deferred final factorization and untouched prior
    n = 10
    Q10 = _tri(n, 2.0, -0.8)
    lik10 = PoissonLikelihood(np.array([2, 1, 3, 0, 4, 1, 2, 3, 1, 0], dtype=float))
    wg = WorkspaceGMRF(np.zeros(n), Q10, **kw())
    Q_copy = Q10.copy()
    res = gaussian_approximation(wg, lik10)
    assert not res.workspace.numeric_valid
    v = res.var()
    assert np.all(v > 0) and v.size == n and res.workspace.numeric_valid and np.isfinite(res.logdetcov())
    assert (wg.precision_matrix() != Q_copy).nnz == 0 and (res.precision_matrix() != Q_copy).nnz > 0


# ---------------------------------------------------------------------------- test/gaussian_approximation/test_predictive_convergence.jl
def test_predictive_convergence_look_ahead_gate():          # :32-58
    from gmrf_b200.workspace_gmrf import _predict_converged as pc
    tol = 1e-8
    assert pc(1e-6, 1e-2, 1.0, 1.0, tol, 2)                  # quadratic contraction, both steps undamped
    assert not pc(1e-6, 1e-2, 1.0, 1.0, tol, 1)              # first iteration: nothing to estimate from
    assert not pc(1e-6, 1e-2, 0.999, 1.0, tol, 2)            # damped step
    assert not pc(1e-6, 1e-2, 1.0, 0.316, tol, 2)            # previous step damped
    assert not pc(1e-4, 1e-2, 1.0, 1.0, tol, 2)              # contraction too slow
    assert not pc(1e-2, 1e-2, 1.0, 1.0, tol, 2)              # no contraction


@pytest.mark.parametrize("kw", BACKENDS)
def test_predictive_convergence_reuses_the_factorization(kw):   # :60-125 (workspace path)
    from gmrf_b200.gmrf import GMRF, gaussian_approximation as ga_gmrf
    n = 200
    Q = _tri(n, 2.01, -1.0)
    y = np.round(np.exp(2.0 * np.sin(np.linspace(0.0, 4.0 * np.pi, n))))
    lik = PoissonLikelihood(y)
    warm = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **dense_kw()), lik).mean()
    Q11 = sp.csc_matrix(1.1 * Q)
    tolkw = dict(newton_dec_tol=1e-8, mean_change_tol=1e-12)
    s_on, s_off = {}, {}
    prior = WorkspaceGMRF(np.zeros(n), Q11, **kw())
    post_on = gaussian_approximation(prior, lik, x0=warm, stats=s_on, **tolkw)
    post_off = gaussian_approximation(prior, lik, x0=warm, predictive_convergence=False, stats=s_off, **tolkw)
    assert s_off["iterations"] > 1 and s_on["iterations"] == s_off["iterations"] - 1     # one loop trip fewer ...
    assert s_on["refactorizations"] == s_off["refactorizations"] - 1                      # ... i.e. one factorization fewer
    assert np.allclose(post_on.mean(), post_off.mean(), rtol=1e-6, atol=1e-10)
    assert np.allclose(post_on.precision.toarray(), post_off.precision.toarray(), rtol=1e-8)

    def returned_decrement(post):                            # g' H^-1 g at the returned mode (:14-18)
        x = post.mean()
        g = Q11 @ x - lik.loggrad(x)
        return float(g @ np.linalg.solve(post.precision.toarray(), g))

    assert returned_decrement(post_on) < 1e-3 * tolkw["newton_dec_tol"]       # the chord steps land far inside
    assert returned_decrement(post_off) < tolkw["newton_dec_tol"]
    # same mode as the cache-backed (plain GMRF) path
    gk = {"backend_type": DenseBackend} if kw is dense_kw else {"device": 0}
    cache_post = ga_gmrf(GMRF(np.zeros(n), Q11, **gk), lik, x0=warm, **tolkw)
    assert np.allclose(post_on.mean(), cache_post.mean(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("kw", BACKENDS)
def test_predictive_convergence_is_inert_while_the_line_search_damps(kw):   # :127-152
    m = 5
    weak = sp.csc_matrix(0.01 * sp.identity(m))
    lik = PoissonLikelihood(np.array([200, 50, 500, 10, 1000], dtype=float))
    tolkw = dict(newton_dec_tol=1e-8, mean_change_tol=1e-12)
    for extra in (dict(step_recovery="sqrt"), dict()):
        on = gaussian_approximation(WorkspaceGMRF(np.zeros(m), weak, **kw()), lik, **extra, **tolkw)
        off = gaussian_approximation(WorkspaceGMRF(np.zeros(m), weak, **kw()), lik, predictive_convergence=False, **extra, **tolkw)
        assert np.array_equal(on.mean(), off.mean())


# ------------------------------------------------------------ remaining behaviours of test_gaussian_approximation.jl on this path
@pytest.mark.parametrize("kw", BACKENDS)
def test_newton_loop_edge_behaviours(kw):
    from gmrf_b200.gmrf import GMRF, gaussian_approximation as ga_gmrf
    gk = {"backend_type": DenseBackend} if kw is dense_kw else {"device": 0}
    # non-convergence path, max_iter = 1 (:238-255): still a finite approximation of the right type
    n = 5
    far = PoissonLikelihood(np.array([10, 15, 8, 12, 20], dtype=float))
    res = ga_gmrf(GMRF(np.zeros(n), sp.identity(n, format="csc"), **gk), far, max_iter=1)
    assert isinstance(res, GMRF) and res.mean().size == n and np.all(np.isfinite(res.mean()))
    # extreme counts under a weak prior (:323-348, :467-): the line search keeps the iterates finite
    one = ga_gmrf(GMRF(np.zeros(1), sp.csc_matrix(np.array([[0.01]])), **gk), PoissonLikelihood(np.array([100.0])))
    assert np.isfinite(one.mean()[0]) and abs(one.mean()[0] - np.log(100.0)) < 1.0 and one.precision[0, 0] > 0
    y5 = np.array([200, 50, 500, 10, 1000], dtype=float)
    many = gaussian_approximation(WorkspaceGMRF(np.zeros(5), sp.csc_matrix(0.01 * sp.identity(5)), **kw()), PoissonLikelihood(y5))
    assert np.all(np.isfinite(many.mean())) and np.all(np.abs(many.mean() - np.log(y5)) < 1.0)
    # convergence criteria agree (:302-321); Gaussian likelihood through the generic loop
    yv = np.array([0.1, -0.1, 0.2, -0.2])
    lik = LinearGaussianLikelihood(sp.identity(4, format="csr"), yv, 0.5)
    r1 = ga_gmrf(GMRF(np.zeros(4), sp.identity(4, format="csc"), **gk), lik, newton_dec_tol=1e-10)
    r2 = ga_gmrf(GMRF(np.zeros(4), sp.identity(4, format="csc"), **gk), lik, mean_change_tol=1e-8)
    assert np.linalg.norm(r1.mean() - r2.mean()) < 1e-6 and np.allclose(r1.mean(), yv * 4.0 / 5.0, rtol=1e-10)
    # :retry_full is the default (:424-430), and a converged mode is a fixed point of the iteration (:450-465)
    n = 200
    Q = _tri(n, 2.01, -1.0)
    lik200 = PoissonLikelihood(np.round(np.exp(3.0 * np.sin(np.linspace(0.0, 6.0 * np.pi, n)))))
    tight = dict(newton_dec_tol=1e-12, mean_change_tol=1e-12)
    a = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik200, max_iter=8, **tight)
    b = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik200, max_iter=8, step_recovery="retry_full", **tight)
    assert np.array_equal(a.mean(), b.mean())
    star = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik200, **tight).mean()
    warm = gaussian_approximation(WorkspaceGMRF(np.zeros(n), Q, **kw()), lik200, x0=star, **tight).mean()
    gn = lambda x: np.max(np.abs(Q @ x - lik200.loggrad(x)))
    assert np.allclose(warm, star, atol=1e-10) and gn(warm) < 1e-12


# ------------------------------------------------------------------------------------------- host logic at realistic sizes, on the CPU
def test_newton_loop_and_theta_loop_at_10k_dofs_on_the_cpu_port():
    """BASELINE config 2 / 3 in the small (2D Matern alpha = 3, 10,201 dofs) through the same host code, with the CPU
    supernodal port as the backend: the Newton loop ends at a stationary point with a handful of numeric-only
    refactorizations, posterior variances are sane, and a theta loop on the shared workspace reproduces SuperLU's
    log-determinants."""
    import scipy.sparse.linalg as spl
    from cpu_port_backend import CpuPortBackend
    coords, cells = spde.mesh2d(100)
    m = spde.MaternSPDE(coords, cells, 1)
    model = MaternModel(m)
    n = m.n
    ws = make_workspace(model, {"backend_type": CpuPortBackend, "ordering": spde.geometric_nd_perm((101, 101), leaf=64, width=3)},
                        tau=1.0, range_=0.3)
    lam = np.exp(0.5 + 0.5 * np.sin(2 * np.pi * coords[:, 0]) * np.cos(2 * np.pi * coords[:, 1]))
    lik = PoissonLikelihood(np.random.default_rng(1).poisson(lam).astype(float))
    prior = evaluate_with_workspace(model, ws, tau=1.0, range_=0.3)
    stats = {}
    before = ws.backend.refactorizations
    post = gaussian_approximation(prior, lik, newton_dec_tol=1e-10, mean_change_tol=1e-10, stats=stats)
    Q = prior.precision
    x = post.mean()
    assert np.max(np.abs(Q @ x - lik.loggrad(x))) <= 1e-8 * abs(Q).max()
    assert 2 <= stats["refactorizations"] <= 12 and ws.backend.refactorizations - before <= stats["refactorizations"] + 1
    sd = post.std()
    assert sd.shape == (n,) and np.all(np.isfinite(sd)) and sd.min() > 0 and sd.max() < prior.std().max()
    for tau, rng_ in ((0.5, 0.2), (2.0, 0.6)):
        d = evaluate_with_workspace(model, ws, tau=tau, range_=rng_)
        lu = spl.splu(sp.csc_matrix(d.precision), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0, options=dict(SymmetricMode=True))
        ld_ref = float(np.sum(np.log(np.abs(lu.U.diagonal()))))
        assert abs(-d.logdetcov() - ld_ref) <= 1e-10 * abs(ld_ref)
