"""Host-side mirror of boundary B: the plain `GMRF` object and `linear_condition`, which in
the reference reach the factorization through a `LinearSolve.LinearCache` instead of a workspace
(src/gmrf.jl:159-223 constructors, :267 logdetcov, :271-296 _rand!, :318-332 var; src/solvers/{selinv,backward_solve,
logdet}.jl dispatch tables; src/arithmetic/condition/linear.jl:46-64). As there, every construction analyses and
factorizes anew (no symbolic reuse guarantee) -- the workspace path is the one for loops. In the Julia integration this is
`julia/B200LinearSolve.jl` (a LinearSolve algorithm in the manner of ext/GaussianMarkovRandomFieldsPardiso.jl)."""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

from .backend import B200Backend, _csc

__all__ = ["GMRF", "linear_condition", "gaussian_approximation"]


class GMRF:
    """GMRF(mean, Q) or GMRF(information=h, Q): N(Q^-1 h, Q^-1) backed by one B200 factorization of Q."""

    def __init__(self, mean=None, Q=None, information=None, ordering=None, device: int = 0, backend_type=B200Backend,
                 linsolve_cache=None):
        Q = _csc(Q).astype(np.float64)
        n = Q.shape[0]
        if Q.shape[0] != Q.shape[1]:
            raise ValueError("size mismatch")
        self.precision = Q
        self._backend_type, self._device = backend_type, device
        if linsolve_cache is not None:                                  # GMRF(x, Q; linsolve_cache = solver), gmrf.jl:159-193:
            self.linsolve_cache = linsolve_cache                        # the cache already holds the factorization of Q
        else:
            kw = {"device": device, "ordering": ordering} if backend_type is B200Backend else {}
            self.linsolve_cache = backend_type(Q, **kw)                 # symbolic analysis + numeric factorization
        if information is not None:                                     # gmrf.jl:195-223: mean = Q \ h
            information = np.asarray(information, dtype=np.float64)
            if information.size != n:
                raise ValueError("size mismatch")
            self.information = information.copy()
            self.mean_ = np.asarray(self.linsolve_cache.backend_solve(self.information))
        else:
            mean = np.asarray(mean, dtype=np.float64)
            if mean.size != n:
                raise ValueError("size mismatch")
            self.mean_ = mean.copy()
            self.information = None

    def __len__(self):
        return self.precision.shape[0]

    def mean(self):
        return self.mean_

    def precision_matrix(self):
        return self.precision

    def information_vector(self):
        return self.information if self.information is not None else self.precision @ self.mean_

    def logdetcov(self):                                                # _logdet_cov_impl, solvers/logdet.jl:27-31
        return -self.linsolve_cache.compute_logdet()

    def var(self):                                                      # _selinv_diag_impl, solvers/selinv.jl:70-73
        return np.array(self.linsolve_cache.get_selinv_diag())

    def std(self):
        return np.sqrt(self.var())

    def rand(self, rng: np.random.Generator, m: int | None = None):    # _backward_solve_impl, solvers/backward_solve.jl:50-53
        n = len(self)
        if m is None:
            return self.linsolve_cache.backend_backward_solve(rng.standard_normal(n)) + self.mean_
        Z = np.asfortranarray(rng.standard_normal((m, n)).T)
        return self.linsolve_cache.backend_backward_solve(Z) + self.mean_[:, None]

    def sqmahal(self, x):                                               # gmrf.jl:94-97
        d = np.asarray(x, dtype=np.float64) - self.mean_
        return float(d @ (self.precision @ d))

    def gradlogpdf(self, x):                                            # gmrf.jl:100
        return -(self.precision @ (np.asarray(x, dtype=np.float64) - self.mean_))

    def logpdf(self, z):
        r = np.asarray(z, dtype=np.float64) - self.mean_
        return float(-0.5 * (r @ (self.precision @ r)) - 0.5 * self.logdetcov() - 0.5 * len(self) * math.log(2.0 * math.pi))


def linear_condition(gmrf: GMRF, A, Q_eps, y, b=None, obs_precision_contrib=None, **gmrf_kwargs) -> GMRF:
    """Posterior of x | y for y = A x + b + eps, eps ~ N(0, Q_eps^-1) by information-vector arithmetic
    (condition/linear.jl:46-64): Q_post = Q + A' Q_eps A, h_post = h + A' Q_eps (y - b); the posterior mean is one solve."""
    Q_prior = gmrf.precision_matrix()
    n = Q_prior.shape[0]
    A = sp.csr_matrix(A, dtype=np.float64)
    m = A.shape[0]
    Q_eps = sp.identity(m, format="csr") * float(Q_eps) if np.isscalar(Q_eps) else sp.csr_matrix(Q_eps, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    b = np.zeros(m) if b is None else np.asarray(b, dtype=np.float64)
    if obs_precision_contrib is None:
        obs_precision_contrib = A.T @ Q_eps @ A
    Q_post = sp.csc_matrix(Q_prior + obs_precision_contrib)
    Q_post.sort_indices()
    h_post = gmrf.information_vector() + A.T @ (Q_eps @ (y - b))
    return GMRF(information=h_post, Q=Q_post, **gmrf_kwargs)


def gaussian_approximation(prior: GMRF, obs_lik, **kwargs) -> GMRF:
    """Gaussian approximation for a plain `GMRF` prior -- the cache-backed Newton loop of
    src/arithmetic/condition/gaussian_approximation.jl:197-229, :428-498: one solver is set up for the pattern of
    `Q_prior - H(x0)` (`_ga_init_solver` / `_ga_resolve_cache`, :88-110), every iterate is a values-only refactorization
    (`_ga_refactor!` -> `_update_linsolve_cache!`, :61-76, :112-116) plus one solve, and the posterior `GMRF` adopts the
    solver (`_ga_make_posterior`, :126-129). On this backend "one solver per pattern" is exactly a workspace, so the loop is
    the workspace Newton loop on a workspace built for the joint pattern with the prior's ordering; the iterates and the
    result are those of `workspace_gmrf.gaussian_approximation`. Keyword arguments as there."""
    from .latent_model_integration import ones_pattern, copy_values_into
    from .workspace import GMRFWorkspace
    from . import workspace_gmrf as wg

    Q_prior = prior.precision_matrix()
    n = Q_prior.shape[0]
    x_init = np.asarray(kwargs.get("x0") if kwargs.get("x0") is not None else prior.mean(), dtype=np.float64)
    H = obs_lik.loghessian(x_init)
    H_sparse = sp.csc_matrix(sp.diags(H)) if isinstance(H, np.ndarray) and H.ndim == 1 else _csc(H)
    joint = _csc(ones_pattern(Q_prior) + ones_pattern(H_sparse))       # storage of Q_prior - H
    copy_values_into(joint, Q_prior)
    bkw = {"backend_type": prior._backend_type}
    if prior._backend_type is B200Backend:
        bkw["device"] = prior._device
        if joint.nnz == Q_prior.nnz:                                    # same pattern: "deepcopy(cache)" = same ordering
            bkw["ordering"] = prior.linsolve_cache.permutation()
    ws = GMRFWorkspace(joint, **bkw)
    wprior = wg.WorkspaceGMRF(prior.mean(), joint, ws)
    post = wg.gaussian_approximation(wprior, obs_lik, **kwargs)
    post.ensure_loaded()
    ws.ensure_numeric()                                                 # the solver holds the factorization of Q_post
    return GMRF(mean=post.mean(), Q=post.precision, backend_type=prior._backend_type, device=prior._device,
                linsolve_cache=ws.backend)
