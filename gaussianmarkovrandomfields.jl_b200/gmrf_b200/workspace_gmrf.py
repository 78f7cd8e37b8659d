"""Host-side mirror of the callers that sit directly on the workspace API:

* `ConstraintInfo`, `WorkspaceGMRF` -- src/workspace/workspace_gmrf.jl:13-56, :88-305 (distribution facade: `mean`,
  `logdetcov`, `var`, `std`, `rand`, `logpdf`, lazy `ensure_loaded!` ownership of a shared workspace);
* `gaussian_approximation(prior::WorkspaceGMRF, obs_lik)` -- src/workspace/gaussian_approximation.jl:191-328 (Fisher
  scoring: per iterate nzval rebuild, numeric-only refactorization, one solve, line search) with the shared line
  search / look-ahead of src/arithmetic/condition/gaussian_approximation.jl:231-409 and `_prior_local`
  (src/latent_models/local_quadratic.jl:138-155);
* `PoissonLikelihood` -- canonical log link, src/observation_models/exponential_family/canonical_implementations.jl
  :18-24 (loglik), :165-171 (loggrad), :265-270 (loghessian).

In the Julia integration none of this exists: the reference's own code runs unchanged on `B200Backend`. The mirror is
host control flow only (every factorization / solve / selected inversion goes through the C-ABI); it lets the parity
tests read like test/workspace/test_workspace_gmrf.jl and test_workspace_gaussian_approximation.jl, and it carries the
one facade extension of SURVEY.md 8f.1: `rand(rng, m)` draws m samples with ONE blocked half-solve (column i = i-th
`randn!` draw) instead of m single-column ones (workspace_gmrf.jl:275-286).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
from scipy.special import gammaln

from .backend import _csc
from .workspace import GMRFWorkspace

__all__ = ["ConstraintInfo", "WorkspaceGMRF", "PoissonLikelihood", "gaussian_approximation"]

LOG_2PI = math.log(2.0 * math.pi)


class ConstraintInfo:
    """Precomputed quantities of the constraint A x = e (workspace_gmrf.jl:13-56)."""

    def __init__(self, ws: GMRFWorkspace, mu, A, e):
        n = ws.dimension()
        A = sp.csr_matrix(A, dtype=np.float64)
        e = np.asarray(e, dtype=np.float64)
        if A.shape[1] != n:
            raise ValueError(f"Constraint matrix size {A.shape} incompatible with workspace size {n}")
        if A.shape[0] != e.size:
            raise ValueError(f"Constraint matrix rows {A.shape[0]} != constraint vector length {e.size}")
        self.matrix = A
        self.vector = e
        # A~^T = Q^-1 A^T via ONE blocked multi-RHS solve (:37)
        self.A_tilde_T = np.asarray(ws.workspace_solve(np.asfortranarray(A.T.toarray()))).reshape(n, -1)
        self.L_c = np.linalg.cholesky(np.asarray(A @ self.A_tilde_T))
        mu = np.asarray(mu, dtype=np.float64)
        residual = A @ mu - e
        self.constrained_mean = mu - self.A_tilde_T @ self._lc_solve(residual)
        resid_e = e - A @ mu
        r = resid_e.size
        logdet_lc = 2.0 * np.sum(np.log(np.diag(self.L_c)))
        AAt = np.asarray((A @ A.T).todense())
        self.log_constraint_correction = (0.5 * (r * LOG_2PI + logdet_lc + resid_e @ self._lc_solve(resid_e))
                                          - 0.5 * np.linalg.slogdet(AAt)[1])

    def _lc_solve(self, v):
        """(A A~^T)^-1 v through its Cholesky factor (`L_c \\ v`)."""
        y = np.linalg.solve(self.L_c, v)
        return np.linalg.solve(self.L_c.T, y)


class WorkspaceGMRF:
    """GMRF backed by a `GMRFWorkspace` (workspace_gmrf.jl:88-305). Owns a snapshot of its precision values and a
    version tag; `ensure_loaded` reloads them into the (possibly shared) workspace when another owner used it."""

    def __init__(self, mean, Q, ws: GMRFWorkspace | None = None, A=None, e=None, precision_logdet=None, **ws_kwargs):
        Q = _csc(Q).astype(np.float64)
        self.mean_ = np.array(mean, dtype=np.float64)
        if ws is None:
            ws = GMRFWorkspace(Q, **ws_kwargs)                       # :131-138
            self.version = self._next_version(ws)
            ws.loaded_version = self.version
        else:
            if not ws._same_pattern(Q):                              # _check_workspace_pattern :189-199
                raise ValueError("Sparsity pattern of Q does not match workspace pattern. "
                                 "WorkspaceGMRF snapshots must share the workspace's colptr/rowval.")
            self.version = self._next_version(ws)
        self.precision = Q.copy()
        self.workspace = ws
        self.precision_logdet = precision_logdet
        self.constraints = None
        if A is not None:
            # load this GMRF's values so the ConstraintInfo solves use the right Q (:173-178)
            ws.Q.data[:] = Q.data
            ws._invalidate()
            ws.loaded_version = self.version
            self.constraints = ConstraintInfo(ws, self.mean_, A, e)

    @staticmethod
    def _next_version(ws):
        v = ws.next_version
        ws.next_version += 1
        return v

    # workspace_gmrf.jl:230-238
    def ensure_loaded(self):
        ws = self.workspace
        if ws.loaded_version != self.version:
            ws.Q.data[:] = self.precision.data
            ws._invalidate()
            ws.loaded_version = self.version

    def has_constraints(self):
        return self.constraints is not None

    def __len__(self):
        return self.precision.shape[0]

    def precision_matrix(self):
        return self.precision

    # :248-250
    def mean(self):
        return self.mean_ if self.constraints is None else self.constraints.constrained_mean

    # :252-258
    def logdetcov(self):
        if self.precision_logdet is not None:
            return -self.precision_logdet
        self.ensure_loaded()
        return self.workspace.logdet_cov()

    # :260-273
    def var(self):
        self.ensure_loaded()
        base = self.workspace.selinv_diag()
        if self.constraints is None:
            return base
        ci = self.constraints
        B_T = np.linalg.solve(ci.L_c, ci.A_tilde_T.T)
        return np.maximum(base - np.sum(B_T * B_T, axis=0), 0.0)

    def std(self):
        return np.sqrt(self.var())

    # :275-286 ; m is the SURVEY 8f.1 extension: one blocked half solve for m draws, same random stream
    def rand(self, rng: np.random.Generator, m: int | None = None):
        self.ensure_loaded()
        n = len(self)
        if m is None:
            x = self.workspace.backward_solve(rng.standard_normal(n)) + self.mean_
            return self._project(x)
        Z = np.asfortranarray(rng.standard_normal((m, n)).T)        # column i = i-th draw of n normals
        X = self.workspace.backward_solve(Z) + self.mean_[:, None]
        return self._project(X)

    def _project(self, x):
        if self.constraints is None:
            return x
        ci = self.constraints
        residual = ci.matrix @ x - (ci.vector if x.ndim == 1 else ci.vector[:, None])
        return x - ci.A_tilde_T @ ci._lc_solve(residual)

    # :288-305
    def logpdf(self, z):
        self.ensure_loaded()
        z = np.asarray(z, dtype=np.float64)
        r = z - self.mean_
        n = len(self)
        val = -0.5 * (r @ (self.precision @ r)) - 0.5 * self.logdetcov() - 0.5 * n * LOG_2PI
        if self.constraints is not None:
            val += self.constraints.log_constraint_correction
        return float(val)


    # generic AbstractGMRF methods, src/gmrf.jl:94-100 (about the mean `mean(d)` returns: the constrained one, if any)
    def sqmahal(self, x):
        d = np.asarray(x, dtype=np.float64) - self.mean()
        return float(d @ (self.precision @ d))

    def gradlogpdf(self, x):
        return -(self.precision @ (np.asarray(x, dtype=np.float64) - self.mean()))


class PoissonLikelihood:
    """Poisson observations with the canonical log link: y_i ~ Poisson(exp(x[indices[i]])). `loghessian` returns the
    diagonal (length n) of the (diagonal) Hessian, the `Diagonal` of canonical_implementations.jl:265-270."""

    def __init__(self, y, indices=None):
        self.y = np.asarray(y, dtype=np.float64)
        self.indices = None if indices is None else np.asarray(indices, dtype=np.int64)
        self._logfact = gammaln(self.y + 1.0)

    def _eta(self, x):
        return x if self.indices is None else x[self.indices]

    def _embed(self, v, n):
        if self.indices is None:
            return v
        out = np.zeros(n)
        np.add.at(out, self.indices, v)
        return out

    def loglik(self, x):
        eta = self._eta(np.asarray(x, dtype=np.float64))
        return float(np.sum(self.y * eta - np.exp(eta) - self._logfact))

    def loggrad(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self._embed(self.y - np.exp(self._eta(x)), x.size)

    def loghessian(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self._embed(-np.exp(self._eta(x)), x.size)


# ------------------------------------------------------------------------------------------------------------------
# gaussian_approximation: workspace Newton loop
# ------------------------------------------------------------------------------------------------------------------
_FROZEN_FINISH_STEPS = 3          # condition/gaussian_approximation.jl:381


def _diagonal_indices(Q: sp.csc_matrix) -> np.ndarray:
    """nzval positions of the diagonal entries (workspace/gaussian_approximation.jl:9-22), vectorised."""
    n = Q.shape[0]
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(Q.indptr))
    idx = np.flatnonzero(Q.indices == cols)
    if idx.size != n or not np.array_equal(cols[idx], np.arange(n)):
        present = np.zeros(n, dtype=bool)
        present[cols[idx]] = True
        raise ValueError(f"workspace Q has no stored diagonal entry in column {int(np.flatnonzero(~present)[0])}")
    return idx.astype(np.int64)


def _prior_local(prior: WorkspaceGMRF, x):
    """(Q, h, energy) of a materialised Gaussian prior (local_quadratic.jl:138-142): never factorizes."""
    Q = prior.precision
    h = Q @ prior.mean()
    return Q, h, 0.5 * (x @ (Q @ x)) - x @ h


def _prior_energy(Q, h, x):
    return 0.5 * (x @ (Q @ x)) - x @ h


def _merit_atol(scale):
    return 64.0 * np.finfo(np.float64).eps * max(scale, 1.0)       # condition/gaussian_approximation.jl:246


def _sparse_hessian_map(Q: sp.csc_matrix, H: sp.csc_matrix) -> np.ndarray:
    """Q.nzval position of every stored entry of H (workspace/gaussian_approximation.jl:31-61); ValueError if H has a
    nonzero outside Q's pattern."""
    n = Q.shape[1]
    qkey = np.repeat(np.arange(n, dtype=np.int64), np.diff(Q.indptr)) * Q.shape[0] + Q.indices
    hcol = np.repeat(np.arange(n, dtype=np.int64), np.diff(H.indptr))
    hkey = hcol * Q.shape[0] + H.indices
    pos = np.searchsorted(qkey, hkey)
    ok = pos < qkey.size
    ok[ok] = qkey[pos[ok]] == hkey[ok]
    if not np.all(ok):
        k = int(np.flatnonzero(~ok)[0])
        raise ValueError(f"Hessian has nonzero at ({int(H.indices[k]) + 1}, {int(hcol[k]) + 1}) which is outside the "
                         "workspace Q sparsity pattern.")
    return pos


def _update_hessian(ws: GMRFWorkspace, H_k, prior_nzval, diag_idx, sparse_hess_map=None):
    """ws.Q := Q_prior - H, invalidate, drop ownership (workspace/gaussian_approximation.jl:96-129). `H_k` is either the
    diagonal of a diagonal Hessian (1-D array, the reference's `Diagonal`) or a sparse matrix whose pattern lies inside the
    workspace's; the index map of the sparse case is built once per Newton loop and returned."""
    if prior_nzval.size != ws.Q.data.size:
        raise ValueError(f"prior precision has {prior_nzval.size} stored entries but the workspace pattern has {ws.Q.data.size}")
    ws.Q.data[:] = prior_nzval
    if isinstance(H_k, np.ndarray) and H_k.ndim == 1:
        ws.Q.data[diag_idx] -= H_k                                   # _subtract_diagonal_hessian! :63-72
    else:
        H_sparse = _csc(H_k)
        if sparse_hess_map is None:
            sparse_hess_map = _sparse_hessian_map(ws.Q, H_sparse)
        np.subtract.at(ws.Q.data, sparse_hess_map, H_sparse.data)    # _subtract_sparse_hessian! :74-83
    ws._invalidate()
    ws.loaded_version = 0
    return sparse_hess_map


def _constrain_step(step, ws, constraints):
    """KKT projection of the Newton step with one blocked multi-RHS solve (:143-149)."""
    if constraints is None:
        return step
    A = constraints.matrix
    A_tilde_T = np.asarray(ws.workspace_solve(np.asfortranarray(A.T.toarray()))).reshape(step.size, -1)
    L_c = np.linalg.cholesky(np.asarray(A @ A_tilde_T))
    rhs = A @ step
    return step - A_tilde_T @ np.linalg.solve(L_c.T, np.linalg.solve(L_c, rhs))


def _line_search(Q_p, h, energy_k, obs_lik, x_k, step, alpha, max_linesearch_iter, newton_dec_tol, retry_full):
    """Backtracking line search on the neg-log-posterior up to a constant (condition/gaussian_approximation.jl:262-318)."""
    def merit(x):
        return _prior_energy(Q_p, h, x) - obs_lik.loglik(x)

    loglik_k = obs_lik.loglik(x_k)
    obj_accept = (energy_k - loglik_k) + _merit_atol(abs(energy_k) + abs(loglik_k))
    if retry_full and alpha < 1.0:
        x_full = x_k - step
        if merit(x_full) <= obj_accept:
            return x_full, 1.0
    accept = False
    x_new = x_k - alpha * step
    for _ in range(max_linesearch_iter):
        candidate = x_k - alpha * step
        if merit(candidate) <= obj_accept:
            x_new, alpha, accept = candidate, math.sqrt(alpha), True
            break
        alpha *= 0.1
        if alpha * np.max(np.abs(step)) < newton_dec_tol / 1000:
            x_new, accept = candidate, True
            break
    if not accept:
        x_new = x_k - alpha * step
    return x_new, alpha


def _predict_converged(dec_k, dec_prev, alpha, alpha_prev, newton_dec_tol, it):
    if not (it > 1 and alpha == 1.0 and alpha_prev == 1.0):        # :361-364
        return False
    return dec_k * (dec_k / dec_prev) ** 2 < newton_dec_tol


def _neg_score(prior, obs_lik, x):
    Q_p, h, _ = _prior_local(prior, x)
    return (Q_p @ x - h) - obs_lik.loggrad(x)


def _frozen_finish(prior, obs_lik, solve_step, x, newton_dec_tol):
    """Chord steps on the factorization already held (condition/gaussian_approximation.jl:387-404)."""
    for j in range(_FROZEN_FINISH_STEPS):
        g = _neg_score(prior, obs_lik, x)
        step = solve_step(g)
        if j == 0 and not (g @ step < newton_dec_tol):
            return None
        x = x - step
    return x


def gaussian_approximation(prior: WorkspaceGMRF, obs_lik, x0=None, max_iter: int = 50, mean_change_tol: float = 1e-4,
                           newton_dec_tol: float = 1e-5, adaptive_stepsize: bool = True, max_linesearch_iter: int = 10,
                           step_recovery: str = "retry_full", predictive_convergence: bool = True, verbose: bool = False,
                           stats: dict | None = None, device_iterates: bool = False) -> WorkspaceGMRF:
    """Workspace-aware Gaussian approximation by Fisher scoring (workspace/gaussian_approximation.jl:191-313). Every
    iterate rebuilds the nzval of `Q_prior - H(x_k)` on the fixed pattern, refactorizes numerically and solves once.
    `stats` (optional dict) receives the iteration / refactorization / solve counts. `device_iterates=True` keeps the
    prior's values resident on the device and forms every iterate there (`set_base_values` /
    `refactorize_minus_diag`), so an iterate uploads n doubles instead of nnz(Q) -- what a method of `_update_hessian!`
    specialised on the B200 backend does in the Julia integration; the iterates are bit-identical."""
    if step_recovery not in ("retry_full", "sqrt"):
        raise ValueError(f"step_recovery must be :retry_full or :sqrt, got :{step_recovery}")
    retry_full = step_recovery == "retry_full"
    ws = prior.workspace
    prior.ensure_loaded()
    constraints = prior.constraints
    x_k = np.array(prior.mean() if x0 is None else x0, dtype=np.float64)
    diag_idx = _diagonal_indices(ws.Q)
    alpha, dec_prev = 1.0, 0.0
    counts = {"iterations": 0, "refactorizations": 0, "solves": 0}
    on_device = device_iterates and hasattr(ws.backend, "refactorize_minus_diag")
    if on_device:
        ws.backend.set_base_values(prior.precision.data)

    def solve(g):
        counts["solves"] += 1
        return _constrain_step(ws.workspace_solve(g), ws, constraints)

    hess_map = [None]                                                # sparse Hessians: index map built once (:257)
    sparse_on_device = [False]

    def build_result(x_final):
        Q_p, _, _ = _prior_local(prior, x_final)
        hess_map[0] = _update_hessian(ws, obs_lik.loghessian(x_final), Q_p.data, diag_idx, hess_map[0])
        Q_post = sp.csc_matrix((ws.Q.data.copy(), ws.Q.indices, ws.Q.indptr), shape=ws.Q.shape)   # _snapshot_Q
        if stats is not None:
            stats.update(counts)
        if constraints is None:
            return WorkspaceGMRF(x_final, Q_post, ws)
        return WorkspaceGMRF(x_final, Q_post, ws, constraints.matrix, constraints.vector)

    for it in range(1, max_iter + 1):
        counts["iterations"] = it
        Q_p, h, energy_k = _prior_local(prior, x_k)
        H_k = obs_lik.loghessian(x_k)
        g_l = obs_lik.loggrad(x_k)
        hess_map[0] = _update_hessian(ws, H_k, Q_p.data, diag_idx, hess_map[0])
        if on_device and isinstance(H_k, np.ndarray) and H_k.ndim == 1:
            ws.backend.refactorize_minus_diag(H_k)          # same values as ws.Q, formed in HBM
            ws.numeric_valid, ws.selinv_valid, ws.logdet_valid = True, False, False
        elif on_device and hasattr(ws.backend, "refactorize_minus_sparse") and not isinstance(H_k, np.ndarray):
            H_sparse = _csc(H_k)
            if not sparse_on_device[0]:                      # positions uploaded once per loop (pattern is fixed, :257)
                ws.backend.set_hessian_pattern(hess_map[0])
                sparse_on_device[0] = True
            ws.backend.refactorize_minus_sparse(H_sparse.data)   # nnz(H) doubles per iterate instead of nnz(Q)
            ws.numeric_valid, ws.selinv_valid, ws.logdet_valid = True, False, False
        else:
            ws.ensure_numeric()
        counts["refactorizations"] += 1
        neg_score = (Q_p @ x_k - h) - g_l
        step = solve(neg_score)
        alpha_prev = alpha
        if adaptive_stepsize:
            x_new, alpha = _line_search(Q_p, h, energy_k, obs_lik, x_k, step, alpha, max_linesearch_iter, newton_dec_tol, retry_full)
        else:
            x_new = x_k - step
        dec = float(neg_score @ step)
        mean_change = float(np.linalg.norm(x_new - x_k))
        mean_change_rel = mean_change / max(float(np.linalg.norm(x_k)), 1e-10)
        if verbose:
            print(f"  Iter {it}: Newton dec = {dec:.3g}, alpha = {alpha:.3f}")
        if dec < newton_dec_tol or mean_change < mean_change_tol or mean_change_rel < mean_change_tol:
            return build_result(x_new)
        if predictive_convergence and _predict_converged(dec, dec_prev, alpha, alpha_prev, newton_dec_tol, it):
            x_final = _frozen_finish(prior, obs_lik, solve, x_new, newton_dec_tol)
            if x_final is not None:
                return build_result(x_final)
        dec_prev = dec
        x_k = x_new
    return build_result(x_k)
