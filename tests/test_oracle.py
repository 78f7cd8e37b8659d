"""CPU tests: the oracle (oracle/gmrf_oracle.c) pinned against the golden dense-LinearAlgebra vectors of the
reference's own deterministic fixtures (tests/golden/fixtures.npz) with the reference's tolerances, plus the CPU
supernodal baseline against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

import oracle
from gmrf_b200 import _lib, spde
from gmrf_b200.backend import _Handle
from gmrf_b200.introspect import Tables

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)
FIX = make_golden.fixtures()


@pytest.mark.parametrize("name", list(FIX))
@pytest.mark.parametrize("order", ["natural", "reverse", "random"])
def test_oracle_against_golden(name, order):
    Q = FIX[name]
    n = Q.shape[0]
    perm = {"natural": None, "reverse": np.arange(n)[::-1].copy(),
            "random": np.random.default_rng(5).permutation(n)}[order]
    F = oracle.OracleFactor(Q, perm)
    assert F.status == 0
    # tolerances of the reference's tests: logdet rtol 1e-10, solve ~ dense \, selinv diag rtol 1e-8, entries 1e-6
    assert abs(F.logdet() - GOLD[name + "/logdet"]) <= 1e-10 * max(1.0, abs(GOLD[name + "/logdet"]))
    x = F.solve(GOLD[name + "/b"])
    assert np.linalg.norm(x - GOLD[name + "/x"]) <= 1e-10 * np.linalg.norm(GOLD[name + "/x"])
    d = F.selinv_diag()
    assert np.max(np.abs(d - GOLD[name + "/diag_inv"]) / np.abs(GOLD[name + "/diag_inv"])) <= 1e-8
    S = F.selinv()
    got = np.asarray(S[GOLD[name + "/inv_rows"], GOLD[name + "/inv_cols"]]).ravel()
    assert np.allclose(got, GOLD[name + "/inv_vals"], rtol=1e-6, atol=0)
    # tr(Q^-1 B) for the golden's deterministic non-symmetric B on Q's pattern (selinv_dot, backend.jl:265-267):
    # entries of the selected inverse to relative 1e-8 => the trace to 1e-8 * sum |terms|
    Qc, cols, bvals = make_golden.dot_matrix_values(Q)
    tr = float(np.dot(np.asarray(S[Qc.indices, cols]).ravel(), bvals))
    assert abs(tr - GOLD[name + "/dot_value"]) <= 1e-8 * GOLD[name + "/dot_scale"]


def test_oracle_sampling_half_solve_covariance():
    # backward_solve maps z ~ N(0, I) to N(0, Q^-1): (P' L^-T)(P' L^-T)' = Q^-1 exactly (test_gmrf_workspace.jl:85-100)
    Q = FIX["rand20"]
    n = Q.shape[0]
    F = oracle.OracleFactor(Q, np.random.default_rng(1).permutation(n))
    M = F.backward_solve(np.eye(n))
    assert np.allclose(M @ M.T, np.linalg.inv(Q.toarray()), rtol=1e-10, atol=1e-14)


def test_oracle_logdet_scaling_and_not_pd():
    Q = FIX["grid_border"]
    n = Q.shape[0]
    F = oracle.OracleFactor(Q)
    ld = F.logdet()
    F.refactorize(2.0 * Q.data)
    assert abs(F.logdet() - (ld + n * np.log(2.0))) <= 1e-9 * abs(ld)      # test_backend_ordering.jl:61-67
    Qb = Q.copy()
    Qb.data[Qb.indptr[3]:Qb.indptr[4]][Qb.indices[Qb.indptr[3]:Qb.indptr[4]] == 3] = -5.0
    assert oracle.OracleFactor(Qb).status > 0


def test_oracle_empty_and_ragged():
    import scipy.sparse as sp
    F = oracle.OracleFactor(sp.csc_matrix((0, 0)))
    assert F.nnzL == 0 and F.logdet() == 0.0
    F = oracle.OracleFactor(sp.identity(5, format="csc") * 4.0)
    assert abs(F.logdet() - 5 * np.log(4.0)) < 1e-14
    with pytest.raises(ValueError):
        F.refactorize(np.ones(3))


@pytest.mark.parametrize("cells", [6, 10])
def test_cpu_supernodal_baseline_matches_oracle(cells):
    from oracle.cpu_baseline import CpuSupernodalCholesky
    model = spde.MaternSPDE(*spde.mesh3d(cells), 0)
    Q = model.precision(1.3, 0.4)
    h = _Handle(Q.shape[0], Q.indptr, Q.indices, None, _lib.ORDER_ND, device=-1)
    T = Tables(h)
    cpu = CpuSupernodalCholesky(T, threads=2)
    cpu.refactorize(Q.data)
    F = oracle.OracleFactor(Q, T.perm)
    assert cpu.status == 0
    assert abs(cpu.logdet - F.logdet()) <= 1e-11 * abs(F.logdet())
    # supernodal sweeps of the CPU baseline against the oracle's solves
    rhs = np.random.default_rng(0).standard_normal(Q.shape[0])
    x, _ = cpu.solve(rhs)
    assert np.linalg.norm(x - F.solve(rhs)) <= 1e-10 * np.linalg.norm(x)
    xh, _ = cpu.solve(rhs, half=True)
    assert np.linalg.norm(xh - F.backward_solve(rhs)) <= 1e-10 * np.linalg.norm(xh)
    # supernodal Takahashi recursion of the CPU baseline against the simplicial one of the oracle
    cpu.selinv()
    d, d_ref = cpu.selinv_diag(), F.selinv_diag()
    assert np.max(np.abs(d - d_ref) / d_ref) <= 1e-10
    cpu.refactorize(2.0 * Q.data)                      # a second round on the same buffers: variances halve
    cpu.selinv()
    assert np.max(np.abs(cpu.selinv_diag() - 0.5 * d_ref) / d_ref) <= 1e-10
    h.close()


@pytest.mark.parametrize("kind,cells,smooth", [("2d", 24, 1), ("3d", 8, 0), ("st", 8, 0)])
def test_oracle_against_superlu_on_spde_inputs(kind, cells, smooth):
    """Independent cross-check at sizes where dense LinearAlgebra is no longer the natural arm: SciPy's SuperLU
    (`splu`, a different code base and a different factorization, LU with its own ordering) on the SPDE inputs the GPU
    parity tests use. Same tolerances as against the golden vectors."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import splu
    if kind == "st":
        Q = spde.AdvectionDiffusionSSM(*spde.mesh2d(cells), nt=5).posterior(np.arange(0, 40, 3), 400.0)
    else:
        mesh = spde.mesh2d(cells) if kind == "2d" else spde.mesh3d(cells)
        Q = spde.MaternSPDE(*mesh, smooth).precision(0.7, 0.45)
    n = Q.shape[0]
    h = _Handle(n, Q.indptr, Q.indices, None, _lib.ORDER_ND, device=-1)
    F = oracle.OracleFactor(Q, h.perm())
    h.close()
    lu = splu(sp.csc_matrix(Q), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    ld_lu = float(np.sum(np.log(np.abs(lu.U.diagonal()))) + np.sum(np.log(np.abs(lu.L.diagonal()))))
    assert abs(F.logdet() - ld_lu) <= 1e-10 * abs(ld_lu)
    rng = np.random.default_rng(9)
    b = rng.standard_normal(n)
    x, x_lu = F.solve(b), lu.solve(b)
    # both are backward stable; they agree to cond(Q) * eps
    assert np.linalg.norm(Q @ x - b) <= 1e-10 * (np.linalg.norm(b) + abs(Q).max() * np.linalg.norm(x))
    assert np.linalg.norm(x - x_lu) <= 1e-7 * np.linalg.norm(x_lu)
    idx = rng.choice(n, 5, replace=False)
    E = np.zeros((n, 5)); E[idx, np.arange(5)] = 1.0
    var_lu = lu.solve(E)[idx, np.arange(5)]
    assert np.max(np.abs(F.selinv_diag()[idx] - var_lu) / var_lu) <= 1e-8
