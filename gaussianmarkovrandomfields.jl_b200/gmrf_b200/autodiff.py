"""Host-side mirror of the reverse-mode rules that sit on the selected inverse:

* `compute_precision_gradient(Qinv, r, ybar)` -- src/autodiff/precision_gradient.jl:137-146 (sparse method):
  d logpdf / dQ = 0.5 * ybar * (Q^-1 - r r') evaluated on the selected inverse's pattern only;
* `logpdf_pullback(x, z, ybar)` -- the pullback of `rrule(logpdf, x::WorkspaceGMRF, z)`,
  src/workspace/autodiff.jl:8-52, including the Rue & Held constraint-correction terms (:24-38);
* `logdetcov_pullback(x, ybar)` -- `rrule(logdetcov, x::WorkspaceGMRF)`, src/workspace/autodiff.jl:54-91:
  Q-bar = -ybar * selinv(ws), never a constraint term.

In the Julia integration these rules are the reference's own code running on `B200Backend` through `selinv(ws)`; the
mirror exists so the parity tests can exercise that consumer of boundary A on both arms.

The extension of this module is the *contracted* form for fixed-pattern hyperparameter models. When
Q(theta) = sum_j c_j(theta) B_j (Matern SPDE precisions: matern_spde.jl:332-356), the chain rule only ever needs
<Q-bar, B_j>; `logdetcov_basis_gradient` / `logpdf_basis_gradient` return exactly those numbers, with the traces
tr(Q^-1 B_j) contracted on the device against the resident value basis (`gmrf_b200_selinv_dot_basis`), so neither
`sparse(Z)` (38 GB at 1 M dofs) nor any pattern ever crosses PCIe.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = ["compute_precision_gradient", "logpdf_pullback", "logdetcov_pullback", "logdetcov_basis_gradient",
           "logpdf_basis_gradient"]


def compute_precision_gradient(Qinv, r, ybar: float):
    """0.5 * ybar * (Qinv - r r') at the stored positions of `Qinv` (precision_gradient.jl:137-146)."""
    Qinv = sp.csc_matrix(Qinv)
    r = np.asarray(r, dtype=np.float64)
    cols = np.repeat(np.arange(Qinv.shape[1]), np.diff(Qinv.indptr))
    vals = (0.5 * ybar) * (Qinv.data - r[Qinv.indices] * r[cols])
    return sp.csc_matrix((vals, Qinv.indices.copy(), Qinv.indptr.copy()), shape=Qinv.shape)


def logpdf_pullback(x, z, ybar: float = 1.0):
    """(mu_bar, Q_bar, z_bar) of `logpdf(x, z)` for a WorkspaceGMRF (src/workspace/autodiff.jl:8-52). `Q_bar` is sparse
    on the selected inverse's pattern; with constraints the correction term makes it dense, as in the reference."""
    z = np.asarray(z, dtype=np.float64)
    mu_base = x.mean_                                   # unconstrained mean (:9)
    Q = x.precision_matrix()
    r = z - mu_base
    x.ensure_loaded()                                   # :16
    Qinv = x.workspace.selinv()
    Qr = Q @ r
    mu_bar = ybar * Qr
    Q_bar = compute_precision_gradient(Qinv, r, ybar)
    if x.has_constraints():                             # :24-38
        ci = x.constraints
        A = ci.matrix
        resid_e = ci.vector - A @ x.mean_
        mu_bar = mu_bar - ybar * (A.T @ ci._lc_solve(resid_e))
        n_c = ci.L_c.shape[0]
        S_inv = ci._lc_solve(np.eye(n_c))
        w = S_inv @ resid_e
        Q_bar = np.asarray(Q_bar.todense()) + (-0.5 * ybar) * (ci.A_tilde_T @ (S_inv - np.outer(w, w)) @ ci.A_tilde_T.T)
    z_bar = ybar * (-Qr)
    return mu_bar, Q_bar, z_bar


def logdetcov_pullback(x, ybar: float = 1.0):
    """Q_bar of `logdetcov(x)` = -ybar * selinv(ws) on the factor's pattern (src/workspace/autodiff.jl:68-91). The same
    with or without constraints: `logdetcov` is the base log-determinant."""
    if ybar == 0:
        return None                                     # ZeroTangent: the selected inversion is skipped (:73)
    x.ensure_loaded()
    return (-ybar) * x.workspace.selinv()


# ---- contracted forms: gradients with respect to the coefficients of a resident value basis -------------------------
def _basis_traces(ws) -> np.ndarray:
    ws.ensure_selinv()
    return np.asarray(ws.backend.selinv_dot_basis(), dtype=np.float64)


def logdetcov_basis_gradient(x, ybar: float = 1.0) -> np.ndarray:
    """d logdetcov / d c_j = <-ybar * Q^-1, B_j> = -ybar * tr(Q^-1 B_j) for Q = sum_j c_j B_j whose value arrays were
    uploaded with `set_value_basis`: the logdetcov pullback contracted against the basis on the device."""
    x.ensure_loaded()
    return (-ybar) * _basis_traces(x.workspace)


def logpdf_basis_gradient(x, z, basis, ybar: float = 1.0) -> np.ndarray:
    """d logpdf(x, z) / d c_j = 0.5 * ybar * (tr(Q^-1 B_j) - r' B_j r) (unconstrained x): `<Q_bar, B_j>` with `Q_bar` of
    `logpdf_pullback`, the traces taken on the device. `basis` is the (nbasis, nnz) value array given to
    `set_value_basis` (the quadratic forms r' B_j r are O(nnz) host work on the workspace pattern)."""
    if x.has_constraints():
        raise NotImplementedError("contracted logpdf gradient is for unconstrained WorkspaceGMRFs")
    z = np.asarray(z, dtype=np.float64)
    r = z - x.mean_
    x.ensure_loaded()
    tr = _basis_traces(x.workspace)
    Q = x.workspace.Q
    cols = np.repeat(np.arange(Q.shape[1]), np.diff(Q.indptr))
    rr = r[Q.indices] * r[cols]
    quad = np.asarray(basis, dtype=np.float64) @ rr
    return (0.5 * ybar) * (tr - quad)
