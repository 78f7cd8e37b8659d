"""A WorkspaceBackend over the CPU supernodal port in oracle/ (TEST infrastructure, like dense_backend.py): lets the
host logic above the backend -- GMRFWorkspace, WorkspaceGMRF, the Newton loop, model(ws; theta...) -- run on CPU at sizes
the dense stand-in cannot reach (10^4 .. 10^5 dofs). Symbolic analysis by the library (analysis-only handle), every
floating-point operation by oracle/supernodal_cpu.c on the host cores. Never used by product code."""
import numpy as np
import scipy.sparse as sp

from gmrf_b200 import _lib
from gmrf_b200.backend import _Handle, ordering_permutation
from gmrf_b200.introspect import Tables
from oracle.cpu_baseline import CpuSupernodalCholesky


class CpuPortBackend:
    def __init__(self, Q, ordering=None, **_):
        Q = sp.csc_matrix(Q)
        Q.sort_indices()
        self.n = Q.shape[0]
        perm = None if ordering is None or isinstance(ordering, str) else ordering_permutation(Q, ordering)
        code = {"nd": _lib.ORDER_ND, "amd": _lib.ORDER_AMD, "natural": _lib.ORDER_NATURAL}.get(ordering, _lib.ORDER_ND) if isinstance(ordering, str) else _lib.ORDER_ND
        self._h = _Handle(self.n, Q.indptr.astype(np.int64), Q.indices.astype(np.int64), perm, code, device=-1)
        self._cpu = CpuSupernodalCholesky(Tables(self._h))
        self.selinv_cache = None
        self.selinv_diag_cache = None
        self.refactorizations = 0
        self.refactorize(Q)

    def refactorize(self, Q):
        self._cpu.refactorize(np.ascontiguousarray(sp.csc_matrix(Q).data, dtype=np.float64))
        self.status = self._cpu.status
        self.selinv_cache = None
        self.selinv_diag_cache = None
        self.refactorizations += 1

    def _cols(self, f, B):
        B = np.asarray(B, dtype=np.float64)
        if B.ndim == 1:
            return f(B)
        return np.asfortranarray(np.column_stack([f(B[:, j]) for j in range(B.shape[1])]))

    def backend_solve(self, rhs):
        return self._cols(lambda b: self._cpu.solve(b)[0], rhs)

    def backend_backward_solve(self, x):
        return self._cols(lambda b: self._cpu.solve(b, half=True)[0], x)

    def compute_logdet(self):
        return self._cpu.logdet

    def compute_selinv(self):
        pass

    def get_selinv_diag(self):
        if self.selinv_diag_cache is None:
            self._cpu.selinv()
            self.selinv_diag_cache = self._cpu.selinv_diag()
        return self.selinv_diag_cache

    def permutation(self):
        return self._h.perm()
