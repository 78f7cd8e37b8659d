"""Boundary B (plain GMRF + linear_condition through a factorization cache), mirrored from test/test_gmrf.jl:64-76 and
test/test_linearsolve_architecture.jl:6-10,61-69 (fixture 5 of SURVEY.md 8c)."""
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
from gmrf_b200.gmrf import GMRF, linear_condition  # noqa: E402


def gpu_kw():
    return {"device": 0}


def dense_kw():
    return {"backend_type": DenseBackend}


BACKENDS = [pytest.param(dense_kw, id="dense-host-logic"), pytest.param(gpu_kw, id="b200", marks=pytest.mark.gpu)]


def llt_fixture(n=10):
    L = sp.diags([np.ones(n), np.full(n - 1, -0.5)], [0, -1]).tocsc()   # test_linearsolve_architecture.jl:6-10
    return sp.csc_matrix(L @ L.T)


@pytest.mark.parametrize("kw", BACKENDS)
def test_gmrf_operations_match_dense(kw):
    Q = llt_fixture()
    n = Q.shape[0]
    mu = np.arange(n, dtype=float) / n
    d = GMRF(mu, Q, **kw())
    Sigma = np.linalg.inv(Q.toarray())
    assert np.allclose(d.var(), np.diag(Sigma), rtol=1e-10)               # test_linearsolve_architecture.jl:61-69
    assert np.allclose(d.std(), np.sqrt(np.diag(Sigma)), rtol=1e-10)
    assert abs(d.logdetcov() + np.linalg.slogdet(Q.toarray())[1]) <= 1e-10
    z = np.linspace(-1, 1, n)
    want = -0.5 * (z - mu) @ Q.toarray() @ (z - mu) - 0.5 * d.logdetcov() - 0.5 * n * np.log(2 * np.pi)
    assert abs(d.logpdf(z) - want) <= 1e-12 * abs(want)
    X = d.rand(np.random.default_rng(0), 40000)
    assert np.allclose(np.cov(X), Sigma, atol=0.08) and np.allclose(X.mean(axis=1), mu, atol=0.05)
    # information-vector constructor: mean = Q \ h (gmrf.jl:195-223)
    h = Q @ mu
    d2 = GMRF(information=h, Q=Q, **kw())
    assert np.allclose(d2.mean(), mu, rtol=1e-10, atol=1e-13)
    with pytest.raises(ValueError):
        GMRF(np.zeros(n + 1), Q, **kw())


@pytest.mark.parametrize("kw", BACKENDS)
def test_linear_condition_matches_closed_form(kw):
    n = 40
    Q = (sp.diags([np.full(n - 1, -1.0), np.full(n, 2.3), np.full(n - 1, -1.0)], [-1, 0, 1])).tocsc()
    mu = np.sin(np.linspace(0, 3, n))
    prior = GMRF(mu, Q, **kw())
    rng = np.random.default_rng(5)
    idx = rng.choice(n, 12, replace=False)
    A = sp.csr_matrix((np.ones(12), (np.arange(12), idx)), shape=(12, n))
    y = rng.standard_normal(12)
    post = linear_condition(prior, A, 4.0, y, b=0.1 * np.ones(12), **kw())
    Qd = Q.toarray() + 4.0 * (A.T @ A).toarray()
    m_ref = np.linalg.solve(Qd, Q @ mu + 4.0 * (A.T @ (y - 0.1)))
    assert np.allclose(post.mean(), m_ref, rtol=1e-10, atol=1e-12)
    assert np.allclose(post.var(), np.diag(np.linalg.inv(Qd)), rtol=1e-9)
    assert np.allclose(post.precision_matrix().toarray(), Qd, rtol=1e-14)
    assert np.all(post.var()[idx] < prior.var()[idx])                      # observed sites tighten
