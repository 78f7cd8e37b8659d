"""Host-side mirror of `GMRFWorkspace` (src/workspace/gmrf_workspace.jl:31-302): the lazy-invalidation state
machine (three validity flags, cached logdet, version counter) around a backend. In the Julia integration this
file does not exist -- the reference's own GMRFWorkspace is used unchanged with `B200Backend`; the mirror lets
the parity tests read like test/workspace/test_gmrf_workspace.jl.
"""
from __future__ import annotations

import queue
from contextlib import contextmanager

import numpy as np
import scipy.sparse as sp

from . import _lib
from .backend import B200Backend, _csc, ordering_permutation

__all__ = ["GMRFWorkspace", "WorkspacePool", "workspace_solve", "backward_solve", "logdet", "selinv", "selinv_diag",
           "selinv_dot", "selinv_extract_at", "update_precision", "update_precision_values", "dimension"]


class GMRFWorkspace:
    def __init__(self, Q, backend_type=B200Backend, **backend_kwargs):
        Q = _csc(Q).astype(np.float64)
        if Q.shape[0] != Q.shape[1]:
            raise ValueError("Q must be square")
        self.Q = Q.copy()
        self.backend = backend_type(self.Q, **backend_kwargs)
        if hasattr(self.backend, "pin_host_buffer"):
            self.backend.pin_host_buffer(self.Q.data)   # ws.Q.nzval is the H2D source of every refactorization
        n = Q.shape[0]
        self.rhs = np.zeros(n)
        self.solution = np.zeros(n)
        self.numeric_valid = True      # gmrf_workspace.jl:76 (just factorized)
        self.selinv_valid = False
        self.logdet_valid = False
        self.logdet_cache = 0.0
        self.next_version = 1
        self.loaded_version = 0

    # gmrf_workspace.jl:91
    def dimension(self):
        return self.Q.shape[0]

    def _same_pattern(self, Qn):
        return (Qn.shape == self.Q.shape and np.array_equal(Qn.indptr, self.Q.indptr)
                and np.array_equal(Qn.indices, self.Q.indices))

    def _invalidate(self):
        self.numeric_valid = False
        self.selinv_valid = False
        self.logdet_valid = False

    # gmrf_workspace.jl:131-143
    def update_precision(self, Q_new):
        Q_new = _csc(Q_new)
        if not self._same_pattern(Q_new):
            raise ValueError("Sparsity pattern mismatch: Q_new has different colptr/rowval. "
                             "GMRFWorkspace requires the same sparsity pattern across updates.")
        self.Q.data[:] = Q_new.data
        self._invalidate()
        self.loaded_version = 0

    # gmrf_workspace.jl:154-165
    def update_precision_values(self, nzval):
        nzval = np.asarray(nzval, dtype=np.float64)
        if nzval.size != self.Q.data.size:
            raise ValueError(f"nzval length {nzval.size} does not match workspace Q nzval length {self.Q.data.size}")
        _lib.host_copy(self.Q.data, nzval)
        self._invalidate()
        self.loaded_version = 0

    # gmrf_workspace.jl:174-182
    def ensure_numeric(self):
        if not self.numeric_valid:
            self.backend.refactorize(self.Q)
            self.numeric_valid = True
            self.selinv_valid = False
            self.logdet_valid = False

    # gmrf_workspace.jl:190-197
    def ensure_selinv(self):
        if not self.selinv_valid:
            self.ensure_numeric()
            self.backend.compute_selinv()
            self.selinv_valid = True

    # gmrf_workspace.jl:207-215
    def workspace_solve(self, b):
        self.ensure_numeric()
        return self.backend.backend_solve(b)

    # gmrf_workspace.jl:222-229
    def logdet(self):
        if not self.logdet_valid:
            self.ensure_numeric()
            self.logdet_cache = self.backend.compute_logdet()
            self.logdet_valid = True
        return self.logdet_cache

    def logdet_cov(self):
        return -self.logdet()

    # gmrf_workspace.jl:250-263
    def selinv(self):
        self.ensure_selinv()
        return self.backend.get_selinv()

    def selinv_diag(self):
        self.ensure_selinv()
        return self.backend.get_selinv_diag()

    # gmrf_workspace.jl:274-292
    def selinv_dot(self, B):
        self.ensure_selinv()
        return self.backend.selinv_dot(B)

    def selinv_extract_at(self, B):
        self.ensure_selinv()
        return self.backend.selinv_extract_at(B)

    # gmrf_workspace.jl:299-302
    def backward_solve(self, x):
        self.ensure_numeric()
        return self.backend.backend_backward_solve(x)


# free-function spellings used by the reference's tests
def dimension(ws): return ws.dimension()
def workspace_solve(ws, b): return ws.workspace_solve(b)
def backward_solve(ws, x): return ws.backward_solve(x)
def logdet(ws): return ws.logdet()
def selinv(ws): return ws.selinv()
def selinv_diag(ws): return ws.selinv_diag()
def selinv_dot(ws, B): return ws.selinv_dot(B)
def selinv_extract_at(ws, B): return ws.selinv_extract_at(B)
def update_precision(ws, Q): return ws.update_precision(Q)
def update_precision_values(ws, nz): return ws.update_precision_values(nz)


class WorkspacePool:
    """Pool of independent workspaces sharing ONE resolved ordering (src/workspace/workspace_pool.jl:42-119):
    `checkout`/`checkin`/`with_workspace`. One workspace per slot; `devices` assigns slots to GPUs round-robin,
    which is how independent hyperparameter evaluations shard across an 8xB200 box. `share_analysis=True` analyses the
    pattern once and creates the other slots from the exported analysis (`gmrf_b200_create_from_analysis`)."""

    def __init__(self, Q, size: int = 1, ordering=None, devices=(0,), share_analysis: bool = False, **kw):
        Q = _csc(Q)
        perm = ordering_permutation(Q, "nd" if ordering is None else ordering)   # resolved ONCE (workspace_pool.jl:55-58)
        self._q = queue.Queue()
        self.workspaces = []
        blob = None
        for i in range(size):
            if blob is not None:
                # not only the permutation: the whole symbolic analysis of the first slot is handed to the others
                ws = GMRFWorkspace(Q, analysis=blob, device=devices[i % len(devices)], **kw)
            else:
                ws = GMRFWorkspace(Q, ordering=perm, device=devices[i % len(devices)], **kw)
                if share_analysis and hasattr(ws.backend, "export_analysis"):
                    blob = ws.backend.export_analysis()
            self.workspaces.append(ws)
            self._q.put(ws)

    def checkout(self):
        return self._q.get()

    def checkin(self, ws):
        self._q.put(ws)

    @contextmanager
    def with_workspace(self):
        ws = self.checkout()
        try:
            yield ws
        finally:
            self.checkin(ws)
