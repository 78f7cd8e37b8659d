"""Host-side mirror of the hyperparameter-loop entry points:

* `make_workspace(model; theta_ref...)`, `GMRFWorkspace(model; ...)`, `GMRFWorkspace(model, obs_lik; ...)` with the joint
  prior + observation-Hessian pattern -- src/workspace/latent_model_integration.jl:32, :100-104, :117-135;
* `(model)(ws; theta...)` / `_evaluate_with_workspace` -- :151-185: fresh mean / precision from the hyperparameters,
  padded into the workspace pattern, loaded with `update_precision!`, returned as a `WorkspaceGMRF` (constraints and the
  `precision_logdet` structure hook carried along);
* `_pad_to_workspace_pattern` -- :214-250, `_copy_values_into!` -- :252-275, `_ones_pattern` -- :196-198;
* `make_workspace_pool` / `WorkspacePool(model; ...)` -- :45-46, src/workspace/workspace_pool.jl:69-76.

A "latent model" here is any object with the reference's `LatentModel` interface spelled as methods:
`precision_matrix(**theta)`, `mean(**theta)`, `constraints(**theta)` (None or `(A, e)`) and optionally
`precision_logdet(**theta)` (src/latent_models/latent_model.jl:115-137). The models themselves are producers of Q and
stay in Julia; tests bring small stand-ins.

B200 extension: a model that also exposes `basis()` (value arrays on its structural pattern) and
`coefficients(**theta)` is evaluated with the values assembled in HBM -- `evaluate_with_workspace(..., on_device=True)`
uploads the basis once per workspace and afterwards moves `nbasis` doubles per theta instead of nnz(Q).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .backend import _csc
from .workspace import GMRFWorkspace, WorkspacePool
from .workspace_gmrf import WorkspaceGMRF

__all__ = ["make_workspace", "make_workspace_pool", "workspace_for", "evaluate_with_workspace", "pad_to_workspace_pattern",
           "copy_values_into", "ones_pattern"]


def _ensure_sparse(Q) -> sp.csc_matrix:
    """`_ensure_sparse` (src/latent_models/combined.jl:291-292): full symmetric CSC with sorted rows."""
    Q = _csc(Q).astype(np.float64)
    return Q


def ones_pattern(A: sp.csc_matrix) -> sp.csc_matrix:
    """A's stored pattern with all-ones values (:196-198): cancellation-free pattern unions."""
    A = _csc(A)
    return sp.csc_matrix((np.ones(A.indices.size), A.indices.copy(), A.indptr.copy()), shape=A.shape)


def _positions_in(dst: sp.csc_matrix, src: sp.csc_matrix, what: str) -> np.ndarray:
    """nzval position in `dst` of every stored entry of `src` (both CSC with sorted rows); ValueError outside."""
    if src.shape != dst.shape:
        raise ValueError(f"Q has size {src.shape} but workspace expects {dst.shape}.")
    n = dst.shape[1]
    dcol = np.repeat(np.arange(n, dtype=np.int64), np.diff(dst.indptr))
    scol = np.repeat(np.arange(n, dtype=np.int64), np.diff(src.indptr))
    dkey = dcol * dst.shape[0] + dst.indices                    # strictly increasing: columns, then sorted rows
    skey = scol * dst.shape[0] + src.indices
    pos = np.searchsorted(dkey, skey)
    ok = (pos < dkey.size)
    ok[ok] = dkey[pos[ok]] == skey[ok]
    if not np.all(ok):
        k = int(np.flatnonzero(~ok)[0])
        raise ValueError(f"{what} has nonzero at ({int(src.indices[k]) + 1}, {int(scol[k]) + 1}) outside the workspace pattern.")
    return pos


def pad_to_workspace_pattern(Q, ws: GMRFWorkspace) -> sp.csc_matrix:
    """`Q` padded into `ws.Q`'s pattern with zeros at the positions only the workspace has (:214-250); `Q` itself when the
    patterns already match; ValueError if `Q` has entries outside the workspace pattern (they would be lost)."""
    Q = _ensure_sparse(Q)
    if Q.shape != ws.Q.shape:
        raise ValueError(f"Q has size {Q.shape} but workspace expects {ws.Q.shape}.")
    if ws._same_pattern(Q):
        return Q
    pos = _positions_in(ws.Q, Q, "Q")
    vals = np.zeros(ws.Q.data.size)
    vals[pos] = Q.data
    return sp.csc_matrix((vals, ws.Q.indices.copy(), ws.Q.indptr.copy()), shape=ws.Q.shape)


def copy_values_into(dst: sp.csc_matrix, src: sp.csc_matrix) -> None:
    """dst.nzval := src at matching positions, zero elsewhere (`_copy_values_into!`, :252-275)."""
    src = _csc(src)            # positions and values must come from the SAME (sorted CSC) object
    pos = _positions_in(dst, src, "src")
    dst.data[:] = 0.0
    dst.data[pos] = src.data


def workspace_for(model, obs_lik=None, backend_kwargs=None, **theta_ref) -> GMRFWorkspace:
    """`GMRFWorkspace(model; theta...)` (:100-104) or, with an observation likelihood, the workspace on the JOINT pattern
    `pattern(Q_prior) U pattern(H_obs)` holding Q_prior's values (:117-135). The Hessian pattern is read off
    `loghessian` at x = 0, as in the reference."""
    Q_prior = _ensure_sparse(model.precision_matrix(**theta_ref))
    kw = backend_kwargs or {}
    if obs_lik is None:
        return GMRFWorkspace(Q_prior, **kw)
    n = Q_prior.shape[0]
    H = obs_lik.loghessian(np.zeros(n))
    H_sparse = sp.csc_matrix(sp.diags(H)) if isinstance(H, np.ndarray) and H.ndim == 1 else _csc(H)   # _ensure_sparse_hessian
    joint = _csc(ones_pattern(Q_prior) + ones_pattern(H_sparse))
    copy_values_into(joint, Q_prior)
    return GMRFWorkspace(joint, **kw)


def make_workspace(model, backend_kwargs=None, **theta_ref) -> GMRFWorkspace:
    """`make_workspace(m; theta_ref...)` (:32): the workspace a hyperparameter loop evaluates `m` on."""
    return workspace_for(model, None, backend_kwargs, **theta_ref)


def make_workspace_pool(model, size: int = 1, devices=(0,), ordering=None, backend_kwargs=None, **theta_ref) -> WorkspacePool:
    """`make_workspace_pool(m; size, theta_ref...)` (:45-46) -> `WorkspacePool(model; ...)`: one workspace per slot, all
    with the same resolved ordering; `devices` spreads the slots over the GPUs of the box."""
    Q = _ensure_sparse(model.precision_matrix(**theta_ref))
    return WorkspacePool(Q, size=size, ordering=ordering, devices=devices, **(backend_kwargs or {}))


def evaluate_with_workspace(model, ws: GMRFWorkspace, on_device: bool = False, **theta) -> WorkspaceGMRF:
    """`model(ws; theta...)` (:151-185). `on_device=True` (models with `basis()` / `coefficients()`): the values are formed in
    HBM from the resident basis and factorized right away; the host copy of the values (the WorkspaceGMRF snapshot that
    `logpdf` contracts with) is the same combination, so both paths hand back identical objects."""
    mu = np.asarray(model.mean(**theta), dtype=np.float64)
    constraint_info = model.constraints(**theta)
    ld = model.precision_logdet(**theta) if hasattr(model, "precision_logdet") else None
    if on_device:
        basis = model.basis()
        if basis.shape[1] != ws.Q.data.size:
            raise ValueError("on_device evaluation needs a workspace on the model's own structural pattern")
        if getattr(ws, "_value_basis_owner", None) is not model:
            ws.backend.set_value_basis(basis)
            ws._value_basis_owner = model
        coeff = np.asarray(model.coefficients(**theta), dtype=np.float64)
        Q_for_ws = sp.csc_matrix((coeff @ basis, ws.Q.indices.copy(), ws.Q.indptr.copy()), shape=ws.Q.shape)
        ws.Q.data[:] = Q_for_ws.data
        ws.backend.refactorize_combination(coeff)
        ws.numeric_valid, ws.selinv_valid, ws.logdet_valid = True, False, False
        ws.loaded_version = 0
    else:
        Q_for_ws = pad_to_workspace_pattern(model.precision_matrix(**theta), ws)
        ws.update_precision(Q_for_ws)
    if constraint_info is None:
        d = WorkspaceGMRF(mu, Q_for_ws, ws, precision_logdet=ld)
    else:
        A, e = constraint_info
        d = WorkspaceGMRF(mu, Q_for_ws, ws, A=A, e=e, precision_logdet=ld)
    if on_device and constraint_info is None:
        # the workspace holds exactly this GMRF's values AND their factorization: it owns the load, so that the first
        # use does not reload + refactorize through the host path (the reference's lazy reload costs nothing there)
        ws.loaded_version = d.version
    return d
