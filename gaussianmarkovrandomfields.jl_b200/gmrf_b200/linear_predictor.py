"""Marginal variances of a linear predictor eta = A x under a workspace-backed posterior: diag(A Sigma A') read at the
observation-local pattern only (src/linear_predictor_marginals.jl:118-165, `_row_diag_AΣAt` workspace method :137-141 and
the per-row contraction :152-165; constraint correction :180-195). Sigma is touched only at the pattern of A'A through
`selinv_extract_at` (backend.jl:275-279 -> gmrf_b200_selinv_extract), never through the materialised `sparse(Z)` -- the
consumer SURVEY.md 8f.3 lists as "next". Host glue: the contraction is O(m * nnz_per_row^2)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .workspace_gmrf import WorkspaceGMRF

__all__ = ["linear_predictor_variances"]


def _row_diag_A_sigma_At(A: sp.csr_matrix, sigma_local: sp.csc_matrix) -> np.ndarray:
    """v[i] = sum_{j,k} A[i,j] A[i,k] Sigma[j,k] with Sigma given on (a superset of) the pattern of A'A."""
    S = sp.csr_matrix(sigma_local)
    out = np.zeros(A.shape[0])
    for i in range(A.shape[0]):
        lo, hi = A.indptr[i], A.indptr[i + 1]
        cols, vals = A.indices[lo:hi], A.data[lo:hi]
        if cols.size:
            out[i] = vals @ (S[cols][:, cols] @ vals)
    return out


def linear_predictor_variances(ga: WorkspaceGMRF, A) -> np.ndarray:
    A = sp.csr_matrix(A, dtype=np.float64)
    if A.shape[1] != len(ga):
        raise ValueError(f"design matrix has {A.shape[1]} columns but the field has {len(ga)} components")
    ga.ensure_loaded()
    be = ga.workspace.backend
    if hasattr(be, "selinv_quadform_rows"):
        # B200 backend: positions of the index pairs looked up on the host, Sigma contracted against them on the device
        # (gmrf_b200_selinv_quadform_rows) -- Sigma never leaves HBM
        ga.workspace.ensure_selinv()
        v = be.selinv_quadform_rows(A)
    else:
        pattern = sp.csc_matrix(A.T @ A)
        pattern.sort_indices()
        sigma_local = ga.workspace.selinv_extract_at(pattern)
        v = _row_diag_A_sigma_At(A, sigma_local)
    ci = ga.constraints
    if ci is not None:                                   # _subtract_constraint_correction! (:189-195)
        M = A @ ci.A_tilde_T
        B_T = np.linalg.solve(ci.L_c, M.T)
        v = np.maximum(v - np.sum(B_T * B_T, axis=0), 0.0)
    return v
