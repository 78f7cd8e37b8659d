"""Regenerates tests/golden/fixtures.npz: dense-LinearAlgebra answers (numpy/LAPACK: slogdet, solve, inv) for the
deterministic fixtures the reference's own tests use. The reference pins its backends
against exactly these dense identities computed at test time (test/workspace/test_gmrf_workspace.jl:26-57,
test_backend_ordering.jl:19-67, test_linearsolve_architecture.jl:61-69); Julia/CHOLMOD cannot run in this image,
so the stored numbers come from LAPACK, not from the reference binary.   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"))
from gmrf_b200 import spde  # noqa: E402


def fixtures():
    n = 10
    L = sp.diags([np.ones(n), -0.5 * np.ones(n - 1)], [0, -1], format="csc")
    return {
        "grid_border": spde.grid_border_fixture(),                      # test_backend_ordering.jl:9-17
        "grid3d_12": spde.grid3d_fixture(12, 12, 12, 0.1),              # benchmarks/benchmarks.jl:160-179
        "tridiag10": spde.tridiag_fixture(10, 2.0, -0.8),               # test_workspace_gaussian_approximation.jl:8-10
        "llt10": sp.csc_matrix(L @ L.T),                                # test_linearsolve_architecture.jl:6-10
        "rand20": spde.random_spd_fixture(20, 0.3, 42),                 # test_gmrf_workspace.jl:8-13 (own RNG)
        "rand400": spde.random_spd_fixture(400, 0.3, 43),               # test_gmrf_workspace.jl:72
    }


def dot_matrix_values(Q):
    """(Q as sorted CSC, column of every stored entry, values of the deterministic non-symmetric B on Q's pattern) used by
    the `dot_value` goldens; tests rebuild B from this formula instead of storing it."""
    Qc = sp.csc_matrix(Q)
    Qc.sort_indices()
    cols = np.repeat(np.arange(Qc.shape[1]), np.diff(Qc.indptr))
    bvals = np.sin(1.0 + 0.37 * np.arange(Qc.nnz)) + 0.25 * np.cos(0.11 * Qc.indices * (cols + 1))
    return Qc, cols, bvals


def main():
    out = {}
    for name, Q in fixtures().items():
        D = Q.toarray()
        n = D.shape[0]
        rng = np.random.default_rng(abs(hash(name)) % 2**31 if False else sum(map(ord, name)))
        b = rng.standard_normal(n)
        Dinv = np.linalg.inv(D)
        out[name + "/logdet"] = np.array(np.linalg.slogdet(D)[1])
        out[name + "/b"] = b
        out[name + "/x"] = np.linalg.solve(D, b)
        out[name + "/diag_inv"] = np.diag(Dinv).copy()
        # a deterministic sample of inverse entries on Q's pattern (full inverse is too large to store for 1728)
        coo = Q.tocoo()
        sel = np.arange(0, coo.nnz, max(1, coo.nnz // 400))
        out[name + "/inv_rows"] = coo.row[sel]
        out[name + "/inv_cols"] = coo.col[sel]
        out[name + "/inv_vals"] = Dinv[coo.row[sel], coo.col[sel]]
        # tr(Q^-1 B) for a deterministic (non-symmetric) B on Q's pattern -- selinv_dot, backend.jl:265-267 -- with the sum
        # of |terms| as the scale of its tolerance; and tr(Q^-1 Q) = n
        Qc, cols, bvals = dot_matrix_values(Q)
        terms = Dinv[Qc.indices, cols] * bvals
        out[name + "/dot_value"] = np.array(terms.sum())
        out[name + "/dot_scale"] = np.array(np.abs(terms).sum())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
