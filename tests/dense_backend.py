"""Dense numpy stand-in for a WorkspaceBackend (test infrastructure only): lets the host logic above the backend --
GMRFWorkspace state machine, WorkspaceGMRF, the workspace Newton loop -- run on CPU in the `-m "not gpu"` suite, and
serves as the independent arm of the GPU parity tests, the way the reference's tests compare the workspace path with
the plain `GMRF` path (test/workspace/test_workspace_gaussian_approximation.jl:8-32)."""
import numpy as np
import scipy.sparse as sp


class DenseBackend:
    def __init__(self, Q, **_):
        self.n = Q.shape[0]
        self.selinv_cache = None
        self.selinv_diag_cache = None
        self.refactorizations = 0
        Qc = sp.csc_matrix(Q)
        self._indices, self._indptr = Qc.indices.copy(), Qc.indptr.copy()
        self.refactorize(Q)

    def refactorize(self, Q):
        A = Q.toarray() if sp.issparse(Q) else np.asarray(Q)
        self.Qd = 0.5 * (A + A.T) if not np.allclose(A, A.T) else A
        self.L = np.linalg.cholesky(self.Qd)
        self.selinv_cache = None
        self.selinv_diag_cache = None
        self.refactorizations += 1

    # device-side Newton iterates (B200Backend.set_base_values / refactorize_minus_diag), dense stand-in
    def set_base_values(self, nzval):
        self._base = np.array(nzval, dtype=np.float64)

    def refactorize_minus_diag(self, hdiag):
        Q = sp.csc_matrix((self._base.copy(), self._indices, self._indptr), shape=(self.n, self.n))
        self.refactorize(Q - sp.diags(np.asarray(hdiag, dtype=np.float64)))

    # device-side value assembly and basis traces (B200Backend.set_value_basis / refactorize_combination /
    # selinv_dot_basis), dense stand-in
    def set_value_basis(self, basis):
        self._basis = np.array(basis, dtype=np.float64)

    def refactorize_combination(self, coeff):
        vals = np.asarray(coeff, dtype=np.float64) @ self._basis
        self.refactorize(sp.csc_matrix((vals, self._indices, self._indptr), shape=(self.n, self.n)))

    def selinv_dot_basis(self):
        S = np.linalg.inv(self.Qd)
        cols = np.repeat(np.arange(self.n), np.diff(self._indptr))
        return self._basis @ S[self._indices, cols]

    def backend_solve(self, rhs):
        rhs = np.asarray(rhs, dtype=np.float64)
        y = np.linalg.solve(self.L, rhs)
        return np.linalg.solve(self.L.T, y)

    def backend_backward_solve(self, x):
        return np.linalg.solve(self.L.T, np.asarray(x, dtype=np.float64))

    def compute_logdet(self):
        return 2.0 * float(np.sum(np.log(np.diag(self.L))))

    def compute_selinv(self):
        pass

    def get_selinv_diag(self):
        if self.selinv_diag_cache is None:
            self.selinv_diag_cache = np.diag(np.linalg.inv(self.Qd)).copy()
        return self.selinv_diag_cache

    def selinv_extract_at(self, B):
        B = sp.csc_matrix(B)
        S = np.linalg.inv(self.Qd)
        cols = np.repeat(np.arange(B.shape[1]), np.diff(B.indptr))
        return sp.csc_matrix((S[B.indices, cols], B.indices.copy(), B.indptr.copy()), shape=B.shape)

    def selinv_dot(self, B):
        B = sp.csc_matrix(B)
        return float(np.dot(self.selinv_extract_at(B).data, B.data))

    def get_selinv(self):
        if self.selinv_cache is None:
            self.selinv_cache = sp.csc_matrix(np.linalg.inv(self.Qd))
        return self.selinv_cache
