"""GPU parity tests through the C-ABI (ctypes -> libgmrf_b200.so), written to read like the reference's
test/workspace/test_gmrf_workspace.jl, test_backend_ordering.jl and test_cliquetrees_backend.jl. The checker is the
CPU oracle and the golden dense-LinearAlgebra vectors; tolerances are the reference's (logdet 1e-10, solve 1e-10,
selinv diag 1e-8, selinv entries 1e-6, cached/bit-identical paths exact)."""
import importlib.util
import os
import threading

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from gmrf_b200 import _lib, spde
from gmrf_b200.backend import B200Backend, NotPositiveDefinite, PinDenseColumns, ordering_permutation
from gmrf_b200.workspace import (GMRFWorkspace, WorkspacePool, backward_solve, dimension, logdet, selinv, selinv_diag,
                                 selinv_dot, selinv_extract_at, update_precision, update_precision_values,
                                 workspace_solve)

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)
FIX = make_golden.fixtures()


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


# ---------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", list(FIX))
@pytest.mark.parametrize("ordering", [None, "amd", "natural"])
def test_golden_fixtures(name, ordering):
    Q = FIX[name]
    ws = GMRFWorkspace(Q, ordering=ordering)
    assert dimension(ws) == Q.shape[0]
    assert abs(logdet(ws) - GOLD[name + "/logdet"]) <= 1e-10 * max(1.0, abs(GOLD[name + "/logdet"]))
    assert _rel(workspace_solve(ws, GOLD[name + "/b"]), GOLD[name + "/x"]) <= 1e-10
    d = selinv_diag(ws)
    assert np.max(np.abs(d - GOLD[name + "/diag_inv"]) / GOLD[name + "/diag_inv"]) <= 1e-8
    S = selinv(ws)
    got = np.asarray(S[GOLD[name + "/inv_rows"], GOLD[name + "/inv_cols"]]).ravel()
    assert np.allclose(got, GOLD[name + "/inv_vals"], rtol=1e-6, atol=0)


# ---------------------------------------------------------------------------------------------- test_gmrf_workspace.jl
class TestGMRFWorkspace:
    n = 20
    Q = FIX["rand20"]
    Qd = FIX["rand20"].toarray()
    Qinv = np.linalg.inv(FIX["rand20"].toarray())

    def test_solve_vector_and_matrix(self):
        ws = GMRFWorkspace(self.Q)
        rng = np.random.default_rng(0)
        b = rng.standard_normal(self.n)
        assert _rel(workspace_solve(ws, b), np.linalg.solve(self.Qd, b)) <= 1e-10
        B = rng.standard_normal((self.n, 13))
        X = workspace_solve(ws, B)
        assert X.shape == B.shape and _rel(X, np.linalg.solve(self.Qd, B)) <= 1e-10

    def test_selinv_dot(self):
        ws = GMRFWorkspace(self.Q)
        assert abs(selinv_dot(ws, self.Q) - self.n) <= 1e-8 * self.n         # tr(Q^-1 Q) = n
        B = self.Q.copy()
        B.data = np.random.default_rng(1).standard_normal(B.nnz)
        S = selinv(ws)
        want = float(S.multiply(B).sum())
        assert abs(selinv_dot(ws, B) - want) <= 1e-10 * abs(want)

    @pytest.mark.parametrize("name", ["rand20", "rand400"])
    def test_selinv_extract_at_bit_identical(self, name):
        Qt = FIX[name]
        ws = GMRFWorkspace(Qt)
        Se = selinv_extract_at(ws, Qt)
        Sf = selinv(ws)
        assert np.array_equal(Se.indptr, Qt.indptr) and np.array_equal(Se.indices, Qt.indices)
        full = np.asarray(Sf[Qt.tocoo().row, Qt.tocoo().col]).ravel()
        assert np.array_equal(Se.tocoo().data, full)                          # bit-identical

    def test_selinv_extract_outside_pattern_is_zero(self):
        # positions outside the (supernodal, relaxed) factor pattern read as 0.0; inside they are Q^-1 entries
        Q = FIX["grid3d_12"]
        n = Q.shape[0]
        ws = GMRFWorkspace(Q)
        rng = np.random.default_rng(4)
        B = sp.random(n, n, density=0.002, random_state=rng, format="csc") + sp.identity(n, format="csc")
        B = sp.csc_matrix(B); B.sort_indices()
        Se = selinv_extract_at(ws, B).tocoo()
        Sf = selinv(ws)
        pat = sp.csc_matrix((np.ones(Sf.nnz), Sf.indices, Sf.indptr), shape=Sf.shape)
        inside = np.asarray(pat[Se.row, Se.col]).ravel() > 0
        assert inside.any() and (~inside).any()
        assert np.all(Se.data[~inside] == 0.0)
        assert np.array_equal(Se.data[inside], np.asarray(Sf[Se.row[inside], Se.col[inside]]).ravel())
        inv = np.linalg.inv(Q.toarray())
        assert np.allclose(Se.data[inside], inv[Se.row[inside], Se.col[inside]], rtol=1e-6)

    def test_backward_solve_is_a_sampler(self):
        ws = GMRFWorkspace(self.Q)
        M = backward_solve(ws, np.eye(self.n))                                # columns = P' L^-T e_i
        assert np.allclose(M @ M.T, self.Qinv, rtol=1e-10, atol=1e-14)        # exact covariance identity
        rng = np.random.default_rng(123)
        Z = rng.standard_normal((self.n, 50000))
        X = backward_solve(ws, Z)
        emp = (X * X).mean(axis=1)
        assert np.allclose(emp, np.diag(self.Qinv), rtol=0.1)                 # test_gmrf_workspace.jl:85-100

    def test_update_precision_and_values(self):
        ws = GMRFWorkspace(self.Q)
        Q2 = self.Q.copy(); Q2.data *= 2.0
        update_precision(ws, Q2)
        b = np.random.default_rng(2).standard_normal(self.n)
        assert _rel(workspace_solve(ws, b), np.linalg.solve(Q2.toarray(), b)) <= 1e-10
        assert abs(logdet(ws) - np.linalg.slogdet(Q2.toarray())[1]) <= 1e-10 * abs(logdet(ws))
        assert np.allclose(selinv_diag(ws), np.diag(np.linalg.inv(Q2.toarray())), rtol=1e-8)
        update_precision_values(ws, self.Q.data * 3.0)
        assert abs(logdet(ws) - np.linalg.slogdet(3.0 * self.Qd)[1]) <= 1e-10 * abs(logdet(ws))

    def test_persistent_buffers_across_many_updates(self):
        ws = GMRFWorkspace(self.Q)
        rng = np.random.default_rng(7)
        for k in range(1, 7):
            Qk = self.Q.copy().tolil()
            Qk = sp.csc_matrix(Qk) * (0.5 + k)
            Qk = Qk + sp.diags(0.3 * k * np.arange(1, self.n + 1) / self.n)
            Qk = sp.csc_matrix(Qk); Qk.sort_indices()
            update_precision(ws, Qk)
            b = rng.standard_normal(self.n)
            D = Qk.toarray()
            assert _rel(workspace_solve(ws, b), np.linalg.solve(D, b)) <= 1e-10
            assert abs(logdet(ws) - np.linalg.slogdet(D)[1]) <= 1e-10 * abs(logdet(ws))
            assert np.allclose(selinv_diag(ws), np.diag(np.linalg.inv(D)), rtol=1e-8)

    def test_pattern_mismatch_error(self):
        ws = GMRFWorkspace(self.Q)
        with pytest.raises(ValueError):
            update_precision(ws, spde.random_spd_fixture(self.n, 0.1, 99))
        with pytest.raises(ValueError):
            update_precision_values(ws, np.ones(3))
        with pytest.raises(ValueError):
            ws.backend.refactorize(np.ones(5))                                 # nnz mismatch inside the backend

    def test_lazy_invalidation_and_caches(self):
        ws = GMRFWorkspace(self.Q)
        d1 = selinv_diag(ws).copy()
        Q2 = self.Q.copy(); Q2.data *= 2.0
        update_precision(ws, Q2)
        d2 = selinv_diag(ws)
        assert np.allclose(d2, 0.5 * d1, rtol=1e-8) and not np.allclose(d1, d2)
        s1, s2 = selinv(ws), selinv(ws)
        assert s1 is s2                                                        # cached object
        assert selinv_diag(ws) is selinv_diag(ws)

    def test_selinv_lazy_materialisation_and_bit_equality(self):
        ws = GMRFWorkspace(self.Q)
        d = selinv_diag(ws)
        assert ws.backend.selinv_cache is None                                 # diagonal-only path builds no CSC
        ws2 = GMRFWorkspace(self.Q)
        S = selinv(ws2)
        assert ws2.backend.selinv_cache is not None
        assert np.array_equal(selinv_diag(ws2), S.diagonal())
        assert np.array_equal(d, S.diagonal())                                 # bit-for-bit (test_gmrf_workspace.jl:220-223)


# ---------------------------------------------------------------------------------------------- test_backend_ordering.jl
def test_backend_ordering_override():
    Q = FIX["grid_border"]
    N = Q.shape[0]
    rhs = np.random.default_rng(0).standard_normal(N)
    ws0 = GMRFWorkspace(Q)
    x0, ld0 = ws0.backend.backend_solve(rhs), ws0.backend.compute_logdet()
    d0 = ws0.selinv_diag()
    for ordering in (np.arange(N)[::-1].copy(), "amd", PinDenseColumns("amd"), PinDenseColumns("nd")):
        ws = GMRFWorkspace(Q, ordering=ordering)
        assert _rel(ws.backend.backend_solve(rhs), x0) <= 1e-10
        assert abs(ws.backend.compute_logdet() - ld0) <= 1e-10 * abs(ld0)
        assert np.allclose(ws.selinv_diag(), d0, rtol=1e-8)
    p = ordering_permutation(Q, PinDenseColumns("amd"))
    assert p[-1] == N - 1
    ws = GMRFWorkspace(Q, ordering="amd")
    update_precision(ws, 2.0 * Q)
    ws.ensure_numeric()
    assert abs(ws.backend.compute_logdet() - (ld0 + N * np.log(2.0))) <= 1e-9 * abs(ld0)


def test_pool_shares_one_ordering_and_threads():
    Q = FIX["grid_border"]
    N = Q.shape[0]
    pool = WorkspacePool(Q, size=4, ordering="amd")
    perms = [ws.backend.permutation() for ws in pool.workspaces]
    assert all(np.array_equal(perms[0], p) for p in perms)
    D = Q.toarray()
    rng = np.random.default_rng(3)
    rhs = rng.standard_normal((N, 20))
    out, errs = [None] * 20, []

    def work(i):                     # 20 tasks over 4 workspaces (test_cliquetrees_backend.jl:77-108)
        try:
            with pool.with_workspace() as ws:
                scale = 1.0 + 0.1 * i
                ws.update_precision_values(Q.data * scale)
                out[i] = (ws.workspace_solve(rhs[:, i]), ws.logdet(), scale)
        except Exception as e:       # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(20)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for i, (x, ld, scale) in enumerate(out):
        assert _rel(x, np.linalg.solve(D * scale, rhs[:, i])) <= 1e-10
        assert abs(ld - np.linalg.slogdet(D * scale)[1]) <= 1e-10 * abs(ld)


# ---------------------------------------------------------------------------------------------- vs the oracle on SPDE inputs
@pytest.mark.parametrize("kind,cells,smooth", [("2d", 40, 1), ("3d", 12, 0), ("3d", 8, 1), ("2d", 64, 0)])
def test_spde_parity_with_oracle(kind, cells, smooth):
    mesh = spde.mesh2d(cells) if kind == "2d" else spde.mesh3d(cells)
    model = spde.MaternSPDE(*mesh, smooth)
    Q = model.precision(0.7, 0.45)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    F = oracle.OracleFactor(Q, be.permutation())
    assert np.array_equal(be.colcounts(), F.colcount)            # nnz(L) / column counts bit-exact under the same ordering
    assert be.info()["nnz_l"] == F.nnzL
    assert abs(be.compute_logdet() - F.logdet()) <= 1e-10 * abs(F.logdet())
    rng = np.random.default_rng(0)
    b = rng.standard_normal((n, 3))
    x = be.backend_solve(b)
    assert np.linalg.norm(Q @ x - b) <= 1e-10 * np.linalg.norm(b) * max(1.0, np.linalg.norm(x) * abs(Q).max() / np.linalg.norm(b) * 1e-3)
    assert _rel(x, F.solve(b)) <= 1e-8
    z = rng.standard_normal(n)
    assert _rel(be.backend_backward_solve(z), F.backward_solve(z)) <= 1e-8
    assert np.max(np.abs(be.get_selinv_diag() - F.selinv_diag()) / F.selinv_diag()) <= 1e-8
    # new hyperparameters, same pattern: refactorize only (Newton / theta loops)
    Q2 = model.precision(2.5, 0.2)
    be.refactorize(Q2)
    F.refactorize(Q2.data)
    assert abs(be.compute_logdet() - F.logdet()) <= 1e-10 * abs(F.logdet())
    assert np.max(np.abs(be.get_selinv_diag() - F.selinv_diag()) / F.selinv_diag()) <= 1e-8
    be.close()


@pytest.mark.parametrize("kind,cells,smooth", [("2d", 48, 1), ("3d", 12, 0)])
def test_wide_rhs_blocks_match_single_column_solves(kind, cells, smooth):
    """More than 8 right-hand sides go through the GEMM sweeps in blocks of 64 columns (sampling, constraints,
    backend_solve(::Matrix), backend.jl:207-209): same answers as the column-at-a-time path of backend.jl:199-205 and
    as the oracle, for full blocks, padded tails and the half solve used for sampling."""
    mesh = spde.mesh2d(cells) if kind == "2d" else spde.mesh3d(cells)
    Q = spde.MaternSPDE(*mesh, smooth).precision(1.3, 0.4)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    F = oracle.OracleFactor(Q, be.permutation())
    rng = np.random.default_rng(11)
    for m in (9, 64, 70, 130):
        Bm = rng.standard_normal((n, m))
        X = be.backend_solve(Bm)
        assert X.shape == (n, m)
        cols = [0, m // 2, m - 1]
        for c in cols:
            xc = be.backend_solve(Bm[:, c])
            assert _rel(X[:, c], xc) <= 1e-11
            assert _rel(X[:, c], F.solve(Bm[:, c])) <= 1e-8
        R = Q @ X - Bm
        assert np.linalg.norm(R) <= 1e-9 * (np.linalg.norm(Bm) + abs(Q).max() * np.linalg.norm(X))
        Zm = rng.standard_normal((n, m))
        Sm = be.backend_backward_solve(Zm)
        for c in cols:
            assert _rel(Sm[:, c], be.backend_backward_solve(Zm[:, c])) <= 1e-11
            assert _rel(Sm[:, c], F.backward_solve(Zm[:, c])) <= 1e-8
    # a padded leading dimension (ld > n) and in-place solve through the raw entry point
    be.close()


def test_device_side_value_assembly_for_theta_loops():
    """nzval(tau, range) of a Matern model is a linear combination of alpha + 1 fixed value arrays on the structural
    pattern: uploading them once lets a hyperparameter evaluation refactorize without moving nzval over PCIe."""
    model = spde.MaternSPDE(*spde.mesh2d(24), 1)
    Q0 = model.precision(1.0, 0.5)
    be = B200Backend(Q0, device=0)
    be.set_value_basis(model.basis())
    b = np.random.default_rng(0).standard_normal(model.n)
    for tau, rng_ in ((0.3, 0.25), (2.0, 0.7), (1.0, 0.5)):
        Q = model.precision(tau, rng_)
        be.refactorize(Q)
        ld, x = be.compute_logdet(), be.backend_solve(b)
        be.refactorize_combination(model.coefficients(tau, rng_))
        assert abs(be.compute_logdet() - ld) <= 1e-12 * abs(ld)
        assert _rel(be.backend_solve(b), x) <= 1e-9            # cond(Q) ~ 1e8 amplifies the 1-ulp value differences
        D = Q.toarray()
        assert abs(be.compute_logdet() - np.linalg.slogdet(D)[1]) <= 1e-10 * abs(ld)
    with pytest.raises(ValueError):
        be.set_value_basis(np.ones((2, 5)))
    be.close()


def test_newton_iterate_formed_on_the_device():
    """Q_k = Q_prior - Diagonal(h_k) with the prior's values resident in HBM must equal the host-assembled iterate
    (_update_hessian!, workspace/gaussian_approximation.jl:96-129) bit for bit: the same subtraction on the same entries."""
    model = spde.MaternSPDE(*spde.mesh2d(20), 1)
    Q = model.precision(1.0, 0.5)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    be.set_base_values(Q.data)
    from gmrf_b200.workspace_gmrf import _diagonal_indices
    diag_idx = _diagonal_indices(Q)
    rng = np.random.default_rng(4)
    b = rng.standard_normal(n)
    for _ in range(3):
        h = -np.exp(rng.standard_normal(n))                       # Poisson: loghessian = -exp(eta) <= 0
        Qk = Q.copy()
        Qk.data[diag_idx] -= h                                    # host path: same pattern, explicit zeros kept
        be.refactorize(Qk)
        ld, x = be.compute_logdet(), be.backend_solve(b)
        be.refactorize_minus_diag(h)
        assert be.compute_logdet() == ld and np.array_equal(be.backend_solve(b), x)
    with pytest.raises(ValueError):
        be.refactorize_minus_diag(np.ones(n + 1))
    be.close()


def test_lanes_factorize_a_sweep_side_by_side():
    """A hyperparameter sweep is many numeric factorizations of one pattern whose outputs are log-determinants: a
    handle with `lanes` = B advances B value sets with the same launches. Every lane must reproduce, bit for bit, the
    log-determinant of the one-at-a-time path; lane 0 stays the backend's factor; a non-SPD lane is reported alone."""
    model = spde.MaternSPDE(*spde.mesh2d(32), 1)
    thetas = [(0.2, 0.15), (1.0, 0.3), (3.0, 0.6), (0.7, 0.9), (5.0, 0.2)]
    Q0 = model.precision(*thetas[0])
    single = B200Backend(Q0, device=0)
    want = []
    for th in thetas:
        single.refactorize(model.precision(*th))
        want.append(single.compute_logdet())
    _lib.set_option("lanes", 6)
    try:
        be = B200Backend(Q0, ordering=single.permutation(), device=0)
    finally:
        _lib.set_option("lanes", 1)
    assert be.lane_capacity() == 6 and single.lane_capacity() == 1
    nz = np.stack([model.values(*th) for th in thetas])
    ld, st = be.refactorize_lanes(nz)
    # identical bits per lane: the same launches advance every lane, so a lane reproduces what the handle computes for
    # that value set alone (a lane handle schedules for throughput -- bulk path instead of the latency-oriented fused
    # chain steps -- so against a single-lane HANDLE the agreement is to rounding, not to the bit)
    own = []
    for th in thetas:
        be.refactorize(model.precision(*th))
        own.append(be.compute_logdet())
    assert np.array_equal(ld, np.array(own)) and not st.any()
    assert np.allclose(ld, want, rtol=1e-12, atol=0)
    want = own
    be.set_value_basis(model.basis())
    ld2, st2 = be.refactorize_combination_lanes(np.stack([model.coefficients(*th) for th in thetas]))
    assert np.allclose(ld2, want, rtol=1e-12) and not st2.any()
    b = np.random.default_rng(0).standard_normal(model.n)
    x = be.backend_solve(b)                                                 # lane 0 is the ordinary factor
    Q = model.precision(*thetas[0])
    assert np.linalg.norm(Q @ x - b) <= 1e-9 * (np.linalg.norm(b) + abs(Q).max() * np.linalg.norm(x))
    assert abs(be.compute_logdet() - want[0]) <= 1e-12 * abs(want[0])
    bad = nz.copy()
    bad[2] = -bad[2]                                                        # lane 2 is not positive definite
    ld3, st3 = be.refactorize_lanes(bad)
    assert st3[2] > 0 and not st3[[0, 1, 3, 4]].any()
    assert np.array_equal(ld3[[0, 1, 3, 4]], np.array(want)[[0, 1, 3, 4]])
    with pytest.raises(ValueError):
        be.refactorize_lanes(np.ones((7, nz.shape[1])))                     # more lanes than the handle holds
    with pytest.raises(ValueError):
        single.refactorize_lanes(nz[:2])
    be.close(); single.close()


def test_spatiotemporal_advection_diffusion_posterior():
    """BASELINE config 5 in miniature: block-tridiagonal space-time precision of the implicit-Euler advection-diffusion
    SPDE (time-major), conditioned on point observations of the first slice; posterior marginal variances by selected
    inversion and a block of posterior samples through the blocked half solve."""
    coords, cells = spde.mesh2d(12)
    model = spde.AdvectionDiffusionSSM(coords, cells, nt=8)
    rng = np.random.default_rng(3)
    obs = rng.choice(model.ns, 40, replace=False)                    # observed vertices of time slice 1
    Q = model.posterior(obs, 1.0 / 0.05 ** 2)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    F = oracle.OracleFactor(Q, be.permutation())
    assert np.array_equal(be.colcounts(), F.colcount)
    assert abs(be.compute_logdet() - F.logdet()) <= 1e-10 * abs(F.logdet())
    var = be.get_selinv_diag()
    assert np.max(np.abs(var - F.selinv_diag()) / F.selinv_diag()) <= 1e-8
    assert np.all(var[obs] < 1.05 * 0.05 ** 2)                       # observed sites are pinned by the data
    Z = np.asfortranarray(rng.standard_normal((1024, n)).T)          # column i = i-th randn! draw
    X = be.backend_backward_solve(Z)                                 # 1024 samples: 16 blocks of 64 columns
    for c in (0, 63, 64, 1023):
        assert _rel(X[:, c], F.backward_solve(Z[:, c])) <= 1e-8
    emp = X.var(axis=1)
    assert np.median(np.abs(emp - var) / var) < 0.06                 # samples matched in distribution
    h = rng.standard_normal(n)
    assert _rel(be.backend_solve(h), F.solve(h)) <= 1e-8             # posterior mean solve
    be.close()


def test_schedule_variants_agree():
    """The scheduling options change the launch lists, never the mathematics: split-K with tiny k-slices (the path the
    top supernodes of the 1 M-dof problem take), a small outer block, the generic selected-inversion route for roots and
    stream launches instead of graphs must all reproduce the default answers and the oracle's."""
    model = spde.MaternSPDE(*spde.mesh3d(12), 0)
    Q = model.precision(0.9, 0.5)
    n = Q.shape[0]
    rng = np.random.default_rng(5)
    b = rng.standard_normal((n, 2))
    base = B200Backend(Q, device=0)
    perm = base.permutation()
    F = oracle.OracleFactor(Q, perm)
    ref = (base.compute_logdet(), base.backend_solve(b), base.get_selinv_diag())
    base.close()
    assert np.max(np.abs(ref[2] - F.selinv_diag()) / F.selinv_diag()) <= 1e-8
    variants = [{"splitk_min_k": 32}, {"splitk_min_k": 16, "outer_block": 128}, {"selinv_fast_root": 0}, {"use_graph": 0},
                {"bwd_row_chunk": 64}, {"wide_rhs_min": 1},
                # the three factorization paths: bulk only / fused chain steps everywhere / one-CTA fronts where they fit
                {"fused_front": 0, "fused_chain": 0}, {"fused_front": 0, "chain_max_tiles": 1000000},
                {"fused_chain": 0}, {"front_smem_kb": 60}, {"fused_front": 0, "chain_max_tiles": 40},
                {"asm_gather": 0}, {"asm_gather": 0, "fused_front": 0, "fused_chain": 0}, {"syrk_gather": 1}, {"level_alap": 0}, {"wide_steps": 1}, {"wide_steps": 2}, {"pdl": 0}, {"pdl": 0, "use_graph": 0}, {"pdl_factor": 1}, {"panel_blocked": 0}, {"panel_blocked": 2}, {"potrf_lookahead": 0}, {"potrf_lookahead": 1, "fused_front": 0, "fused_chain": 0}, {"panel_blocked": 2, "fused_front": 0, "chain_max_tiles": 1000000}, {"pdl_multi": 0, "wide_rhs_min": 1}, {"pdl_factor": 1, "use_graph": 0, "fused_front": 0, "fused_chain": 0, "asm_gather": 0},
                {"syrk_gather": 1, "fused_front": 0, "fused_chain": 0}]
    defaults = {"splitk_min_k": 128, "outer_block": 256, "selinv_fast_root": 1, "use_graph": 1, "bwd_row_chunk": 2048,
                "wide_rhs_min": 8, "fused_front": 1, "fused_chain": 1, "chain_max_tiles": 160, "front_smem_kb": 200, "asm_gather": 1, "syrk_gather": 0, "level_alap": 1, "wide_steps": 0, "pdl": 1, "pdl_factor": 0, "pdl_multi": 1, "panel_blocked": 1, "potrf_lookahead": 1}
    try:
        for v in variants:
            for k, val in {**defaults, **v}.items():
                _lib.set_option(k, val)
            be = B200Backend(Q, ordering=perm, device=0)
            ld, x, d = be.compute_logdet(), be.backend_solve(b), be.get_selinv_diag()
            assert abs(ld - ref[0]) <= 1e-12 * abs(ref[0]), v
            assert _rel(x, ref[1]) <= 1e-10, v
            assert np.max(np.abs(d - ref[2]) / ref[2]) <= 1e-10, v
            be.close()
    finally:
        for k, val in defaults.items():
            _lib.set_option(k, val)


def test_wide_solve_steps_agree():
    """Few-RHS triangular solves on supernodes with more than 256 columns: the 64-column steps (default), the 256-column
    steps with their own diagonal-block launches (wide_steps = 1) and the 256-column steps whose head CTA solves the next
    diagonal block in the same launch (wide_steps = 2) are three schedules of the same sweeps. 23^3 vertices: the root
    separator has 529 columns = two full 256-column blocks and a ragged one of 17."""
    model = spde.MaternSPDE(*spde.mesh3d(22), 0)
    Q = model.precision(1.1, 0.4)
    n = Q.shape[0]
    rng = np.random.default_rng(11)
    B = rng.standard_normal((n, 8))
    perm = spde.geometric_nd_perm((23, 23, 23), leaf=64, width=2)
    res = {}
    try:
        for mode in (0, 1, 2, 3, 4):            # 3: 64-column steps WITHOUT programmatic stream serialization
            _lib.set_option("wide_steps", mode % 3 if mode < 4 else 0)      # 4: programmatic launches on the stream, no graph
            _lib.set_option("pdl", 0 if mode == 3 else 1)
            _lib.set_option("use_graph", 0 if mode == 4 else 1)
            be = B200Backend(Q, ordering=perm, device=0)
            assert be.info()["max_ns"] > 512
            out = []
            for m in (1, 2, 3, 8):
                X = be.backend_solve(np.asfortranarray(B[:, :m]) if m > 1 else B[:, 0])
                X = X.reshape(n, -1)
                assert np.max(np.abs(Q @ X - B[:, :m])) <= 1e-10 * np.max(np.abs(B)), (mode, m)
                out.append(X)
            out.append(be.backend_backward_solve(B[:, 0]))
            res[mode] = out
            be.close()
    finally:
        _lib.set_option("wide_steps", 0)
        _lib.set_option("pdl", 1)
        _lib.set_option("use_graph", 1)
    for mode in (1, 2):
        for a, c in zip(res[0], res[mode]):
            assert _rel(c, a) <= 1e-11, mode
    for other in (3, 4):
        for a, c in zip(res[0], res[other]):  # same kernels, same order of operations: identical bits
            assert np.array_equal(a, c), other


def test_run_to_run_bit_reproducible():
    model = spde.MaternSPDE(*spde.mesh3d(10), 0)
    Q = model.precision(1.0, 0.3)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(Q.shape[0])
    res = []
    for _ in range(2):
        be = B200Backend(Q, device=0)
        for _ in range(2):
            be.refactorize(Q)
        res.append((be.compute_logdet(), be.backend_solve(b), be.get_selinv_diag().copy(), be.backend_backward_solve(b)))
        be.close()
    assert res[0][0] == res[1][0]
    for a, c in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(a, c)


def test_not_positive_definite_is_reported():
    Q = FIX["grid_border"].copy()
    be = B200Backend(Q, device=0)
    assert be.status == 0
    Qb = Q.copy()
    Qb.data[:] = -Q.data
    be.refactorize(Qb)                       # silent like `check=false` (backend.jl:184) ...
    assert be.status > 0                      # ... but the failing column is reported
    strict = B200Backend(Q, device=0, check=True)
    with pytest.raises(NotPositiveDefinite):
        strict.refactorize(Qb)
    be.refactorize(Q)                        # the handle recovers on the next valid refactorization
    assert be.status == 0 and abs(be.compute_logdet() - GOLD["grid_border/logdet"]) <= 1e-10 * abs(GOLD["grid_border/logdet"])


def test_edge_cases():
    be = B200Backend(sp.csc_matrix(np.array([[4.0]])), device=0)
    assert abs(be.compute_logdet() - np.log(4.0)) < 1e-15
    assert np.allclose(be.backend_solve(np.array([2.0])), [0.5])
    assert np.allclose(be.get_selinv_diag(), [0.25])
    D = sp.identity(37, format="csc") * 2.0
    be = B200Backend(D, device=0)
    assert abs(be.compute_logdet() - 37 * np.log(2.0)) < 1e-13
    assert np.allclose(be.backend_backward_solve(np.ones(37)), np.ones(37) / np.sqrt(2.0))
    with pytest.raises(ValueError):
        be.backend_solve(np.ones(5))
    with pytest.raises(ValueError):
        B200Backend(sp.csc_matrix(np.ones((2, 3))), device=0)


def test_full_size_config1_properties():
    """BASELINE config 1 size (50,625 dofs): size-independent properties instead of the (slow) oracle."""
    model = spde.MaternSPDE(*spde.mesh2d(224), 1)
    Q = model.precision(1.0, 0.3)
    n = Q.shape[0]
    ws = GMRFWorkspace(Q)
    ld = ws.logdet()
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    x = ws.workspace_solve(b)
    # normwise backward error (cond(Q) ~ 1e8 here; the residual relative to |b| alone is conditioning-limited)
    assert np.linalg.norm(Q @ x - b) <= 1e-10 * (np.linalg.norm(b) + abs(Q).sum(axis=1).max() * np.linalg.norm(x))
    idx = rng.choice(n, 5, replace=False)
    E = np.zeros((n, 5)); E[idx, np.arange(5)] = 1.0
    ref = ws.workspace_solve(E)[idx, np.arange(5)]
    assert np.allclose(ws.selinv_diag()[idx], ref, rtol=1e-8)
    update_precision_values(ws, 2.0 * Q.data)
    assert abs(ws.logdet() - (ld + n * np.log(2.0))) <= 1e-10 * abs(ld)
    z = rng.standard_normal(n)
    s = ws.backward_solve(z)                  # x = P' L^-T z  =>  x' (2Q) x = z' z
    assert abs(s @ (2.0 * (Q @ s)) - z @ z) <= 1e-9 * (z @ z)


def test_selinv_tables_survive_an_allocation_failure():
    """build_selinv_tables must be failure-atomic. Inject an allocation failure into each of its steps in
    turn: the call must fail cleanly (no stale Z array, no half-built plan), and the next call must rebuild from scratch
    and give the right variances."""
    Q = FIX["grid3d_12"]
    be = B200Backend(Q, device=0)
    want = np.diag(np.linalg.inv(Q.toarray()))
    try:
        for k in range(1, 8):
            _lib.set_option("debug_alloc_fail_after", k)
            try:
                d = be.get_selinv_diag()
            except Exception as e:                      # the injected failure surfaced as an error, as it must
                assert "injected" in str(e) or "cudaMalloc" in str(e)
                be.selinv_diag_cache = None
                continue
            finally:
                _lib.set_option("debug_alloc_fail_after", 0)
            assert np.allclose(d, want, rtol=1e-8)      # the countdown outlived the table build: a complete, valid result
            break
        _lib.set_option("debug_alloc_fail_after", 0)
        be.selinv_diag_cache = None
        be.refactorize(Q)
        assert np.allclose(be.get_selinv_diag(), want, rtol=1e-8)
    finally:
        _lib.set_option("debug_alloc_fail_after", 0)
    be.close()


def test_adopting_a_factor_from_a_different_analysis_is_refused():
    """a peer created with another ordering must not silently solve with foreign panels; the sender's pivot
    status travels with the factor."""
    Q = FIX["grid3d_12"]
    a = B200Backend(Q, ordering="nd", device=0)
    b = B200Backend(Q, ordering="amd", device=0)
    twin = B200Backend(Q, ordering=a.permutation(), device=0, factorize=False)
    assert a.analysis_fingerprint() != b.analysis_fingerprint()
    assert a.analysis_fingerprint() == twin.analysis_fingerprint()
    with pytest.raises(ValueError):
        b.adopt_factor(a.compute_logdet(), fingerprint=a.analysis_fingerprint(), status=0)
    import torch
    for w in (0, 1):                                   # same analysis: move the arrays device-to-device and adopt
        (ps, ns), (pd, nd) = a.device_array(w), twin.device_array(w)
        assert ns == nd
        torch.cuda.synchronize()
        from gmrf_b200.sharding import _DeviceArray
        torch.as_tensor(_DeviceArray(pd, nd), device="cuda:0").copy_(torch.as_tensor(_DeviceArray(ps, ns), device="cuda:0"))
    torch.cuda.synchronize()
    twin.adopt_factor(a.compute_logdet(), fingerprint=a.analysis_fingerprint(), status=0)
    rhs = np.random.default_rng(1).standard_normal(Q.shape[0])
    assert np.array_equal(twin.backend_solve(rhs), a.backend_solve(rhs)) and twin.compute_logdet() == a.compute_logdet()
    twin.adopt_factor(a.compute_logdet(), fingerprint=a.analysis_fingerprint(), status=17)
    assert twin.status == 17
    for x in (a, b, twin):
        x.close()


def test_linear_predictor_rows_contracted_on_the_device():
    """diag(A Sigma A') (`_row_diag_AΣAt`, src/linear_predictor_marginals.jl:137-165) through gmrf_b200_selinv_quadform_rows:
    the index pairs of every row are looked up on the host, Sigma is contracted against them in HBM. Rows whose pairs all
    lie inside the factor's pattern (point and local-stencil observations) must match the dense answer to the marginal-
    variance tolerance; 0- and 1-based entry give identical bits; an empty row gives 0."""
    from gmrf_b200.backend import index_base
    Q = FIX["grid3d_12"]
    n = Q.shape[0]
    Sigma = np.linalg.inv(Q.toarray())
    rng = np.random.default_rng(8)
    rows = []
    for i in range(40):
        if i == 7:
            rows.append(np.zeros(n))                                  # an empty row
            continue
        r = np.zeros(n)
        j = rng.integers(n)
        nb = Q[:, [j]].nonzero()[0]                                   # j and its stencil neighbours: pairs inside pattern(Q^2) at most
        pick = rng.choice(nb, size=min(len(nb), rng.integers(1, 4)), replace=False)
        r[pick] = rng.standard_normal(pick.size)
        rows.append(r)
    A = sp.csr_matrix(np.array(rows))
    be = B200Backend(Q, device=0)
    v = be.selinv_quadform_rows(A)
    Z = be.get_selinv()
    inside = np.array([all(Z[a, b] != 0.0 for a in A[i].indices for b in A[i].indices) for i in range(A.shape[0])])
    want = np.einsum("ij,jk,ik->i", A.toarray(), Sigma, A.toarray())
    assert inside.sum() >= 20 and v[7] == 0.0
    assert np.allclose(v[inside], want[inside], rtol=1e-8, atol=0)
    # every row equals the contraction of the materialised selected inverse (zeros outside the pattern), whatever the row
    ref = np.array([A[i].data @ (Z[A[i].indices][:, A[i].indices].toarray() @ A[i].data) if A[i].nnz else 0.0 for i in range(A.shape[0])])
    assert np.allclose(v, ref, rtol=1e-12, atol=1e-300)
    with index_base(1):
        be1 = B200Backend(Q, ordering=be.permutation(), device=0)
        assert np.array_equal(be1.selinv_quadform_rows(A), v)
        be1.close()
    with pytest.raises(ValueError):
        be.selinv_quadform_rows(sp.csr_matrix((3, n + 1)))
    be.close()


def test_sparse_hessian_iterate_formed_on_the_device():
    """Q_k = Q_prior - H_k for a SPARSE observation Hessian (`_subtract_sparse_hessian!`, gaussian_approximation.jl:74-83)
    formed in HBM from nnz(H) uploaded values must equal the host-assembled iterate bit for bit."""
    from gmrf_b200.workspace_gmrf import _sparse_hessian_map
    model = spde.MaternSPDE(*spde.mesh2d(20), 1)
    Q = model.precision(1.0, 0.5)
    n = Q.shape[0]
    rng = np.random.default_rng(9)
    # a Hessian pattern inside Q's: the diagonal plus a random third of the off-diagonal entries, symmetric
    M = sp.triu(Q, 1).tocoo()
    keep = rng.random(M.nnz) < 0.33
    U = sp.csc_matrix((np.ones(keep.sum()), (M.row[keep], M.col[keep])), shape=(n, n))
    P = sp.csc_matrix(U + U.T + sp.identity(n)); P.sort_indices()
    be = B200Backend(Q, device=0)
    be.set_base_values(Q.data)
    hmap = _sparse_hessian_map(Q, P)
    be.set_hessian_pattern(hmap)
    b = rng.standard_normal(n)
    for _ in range(3):
        H = P.copy()
        H.data = -0.05 * np.abs(rng.standard_normal(H.nnz))          # negative semi-definite-ish: Q - H stays SPD here
        H = sp.csc_matrix((H + H.T) * 0.5); H.sort_indices()
        assert np.array_equal(H.indices, P.indices)
        Qk = Q.copy()
        np.subtract.at(Qk.data, hmap, H.data)
        be.refactorize(Qk)
        ld, x = be.compute_logdet(), be.backend_solve(b)
        be.refactorize_minus_sparse(H.data)
        assert be.status == 0 and be.compute_logdet() == ld and np.array_equal(be.backend_solve(b), x)
    with pytest.raises(ValueError):
        be.set_hessian_pattern(np.array([0, 0]))                      # duplicates have no single owner
    with pytest.raises(ValueError):
        be.refactorize_minus_sparse(np.ones(3))
    be.close()


@pytest.mark.parametrize("opts", [{}, {"fused_front": 0}, {"fused_front": 0, "fused_chain": 0}, {"fused_front": 0, "asm_gather": 0},
                                  {"fused_front": 0, "syrk_gather": 1}])
def test_supernode_with_many_children(opts):
    """A root with 14 children (more than the 4 whose inverse maps the fused kernels keep resident at once): 14 dense
    blocks of different sizes coupled only through a 9-variable border. Exercises the multi-pass paths of the one-CTA front
    kernel, of the gather extend-add and of the gathering SYRK tail against dense LinearAlgebra."""
    rng = np.random.default_rng(21)
    sizes = [5, 9, 17, 33, 12, 7, 21, 40, 6, 14, 27, 8, 19, 11]
    nb = 9
    n = sum(sizes) + nb
    D = np.zeros((n, n))
    o = 0
    for m in sizes:
        B = rng.standard_normal((m, m))
        D[o:o + m, o:o + m] = B @ B.T + m * np.eye(m)
        C = 0.3 * rng.standard_normal((m, nb)) * (rng.random((m, nb)) < 0.7)
        D[o:o + m, n - nb:] = C
        D[n - nb:, o:o + m] = C.T
        o += m
    Bb = rng.standard_normal((nb, nb))
    D[n - nb:, n - nb:] = Bb @ Bb.T + 40.0 * np.eye(nb)
    Q = sp.csc_matrix(D); Q.sort_indices()
    perm = np.arange(n)                                       # natural order: blocks first, border last -> one root, 14 children
    defaults = {"fused_front": 1, "fused_chain": 1, "asm_gather": 1, "syrk_gather": 0}
    try:
        for k, v in {**defaults, **opts}.items():
            _lib.set_option(k, v)
        be = B200Backend(Q, ordering=perm, device=0)
    finally:
        for k, v in defaults.items():
            _lib.set_option(k, v)
    assert abs(be.compute_logdet() - np.linalg.slogdet(D)[1]) <= 1e-10 * abs(np.linalg.slogdet(D)[1])
    b = rng.standard_normal(n)
    assert _rel(be.backend_solve(b), np.linalg.solve(D, b)) <= 1e-10
    assert np.allclose(be.get_selinv_diag(), np.diag(np.linalg.inv(D)), rtol=1e-8)
    be.close()
