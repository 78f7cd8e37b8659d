"""The 1-based entry path the Julia glue takes (VERDICT r01 missing #4): `SparseMatrixCSC{Float64,Int}` index arrays and a
1-based `perm` go into `gmrf_b200_create`, and `get_perm`, `selinv_pattern`, `factor_pattern`, `selinv_extract`,
`selinv_dot` exchange 1-based indices (julia/B200Backend.jl:59-66,140,156,168,179). The binding shifts the arrays exactly
like that under `index_base(1)`; every result must be IDENTICAL (bitwise) to the 0-based path and meet the reference's
tolerances on the golden fixtures (tests/golden/make_golden.py)."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

from gmrf_b200 import spde
from gmrf_b200.backend import B200Backend, index_base

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)
FIX = make_golden.fixtures()


def _everything(Q, ordering, B):
    be = B200Backend(Q, ordering=ordering, device=0)
    rng = np.random.default_rng(5)
    b = rng.standard_normal((Q.shape[0], 3))
    out = {"perm": be.permutation(), "logdet": be.compute_logdet(), "x": be.backend_solve(b), "half": be.backend_backward_solve(b[:, 0]),
           "diag": be.get_selinv_diag().copy(), "S": be.get_selinv().copy(), "R": be.cholesky_sqrt(),
           "extract": be.selinv_extract_at(B), "dot": be.selinv_dot(B), "colcounts": be.colcounts()}
    be.close()
    return out


@pytest.mark.parametrize("name", list(FIX))
@pytest.mark.parametrize("ordering", [None, "amd", "natural", "user"])
def test_one_based_path_is_identical_and_meets_the_golden_tolerances(name, ordering):
    Q = FIX[name]
    n = Q.shape[0]
    if ordering == "user":
        ordering = np.random.default_rng(n).permutation(n)            # a caller-supplied permutation (shifted by the binding)
    B = Q.copy(); B.data = np.random.default_rng(1).standard_normal(B.nnz)
    with index_base(0):
        r0 = _everything(Q, ordering, B)
    with index_base(1):
        r1 = _everything(Q, ordering, B)
    assert np.array_equal(r0["perm"], r1["perm"]) and np.array_equal(np.sort(r1["perm"]), np.arange(n))
    assert np.array_equal(r0["colcounts"], r1["colcounts"])
    assert r0["logdet"] == r1["logdet"] and r0["dot"] == r1["dot"]
    for k in ("x", "half", "diag"):
        assert np.array_equal(r0[k], r1[k]), k
    for k in ("S", "R", "extract"):
        a, b = r0[k], r1[k]
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data), k
    # ... and the 1-based results against the dense-LinearAlgebra golden vectors
    assert abs(r1["logdet"] - GOLD[name + "/logdet"]) <= 1e-10 * max(1.0, abs(GOLD[name + "/logdet"]))
    assert np.max(np.abs(r1["diag"] - GOLD[name + "/diag_inv"]) / GOLD[name + "/diag_inv"]) <= 1e-8
    got = np.asarray(r1["S"][GOLD[name + "/inv_rows"], GOLD[name + "/inv_cols"]]).ravel()
    assert np.allclose(got, GOLD[name + "/inv_vals"], rtol=1e-6, atol=0)
    R = r1["R"]
    assert abs(R @ R.T - Q).max() <= 1e-10 * abs(Q).max()                 # P'L is a square root of Q
    Sd = r1["S"]
    assert abs(r1["dot"] - float(Sd.multiply(B).sum())) <= 1e-10 * float(abs(Sd.multiply(B)).sum())


def test_one_based_path_on_an_spde_matrix_with_analysis_blob():
    model = spde.MaternSPDE(*spde.mesh2d(40), 1)
    Q = model.precision(0.7, 0.45)
    b = np.random.default_rng(0).standard_normal(Q.shape[0])
    with index_base(1):
        be = B200Backend(Q, device=0)
        blob = be.export_analysis()
        x1, ld1, p1 = be.backend_solve(b), be.compute_logdet(), be.permutation()
        again = B200Backend(Q, analysis=blob, device=0)                   # create_from_analysis with 1-based arrays
        assert np.array_equal(again.backend_solve(b), x1) and again.compute_logdet() == ld1
        again.close(); be.close()
    with index_base(0):
        be0 = B200Backend(Q, ordering=p1, device=0)
        assert np.array_equal(be0.backend_solve(b), x1) and be0.compute_logdet() == ld1
        be0.close()
    assert np.linalg.norm(Q @ x1 - b) <= 1e-10 * (np.linalg.norm(b) + abs(Q).sum(axis=1).max() * np.linalg.norm(x1))
