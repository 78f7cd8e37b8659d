"""Parity on the code paths the benchmark times: inputs large enough that the 128 x 64-tile GEMM,
natural split-K (k >= 1024 slices) and the triangular root route of the selected inversion are actually scheduled --
asserted through the plan introspection counters of gmrf_b200_info -- checked against the CPU supernodal port
(oracle/cpu_baseline.py, itself pinned to the simplicial oracle in tests/test_oracle.py), the oracle's independent
symbolic column counts, and size-independent identities (test/workspace/test_backend_ordering.jl:61-67) at the sizes of
BASELINE configs 2, 3 and 5. Tolerances are north_star's: column counts exact, logdet 1e-10, solves 1e-8 against the
port + normwise backward error 1e-10, marginal variances 1e-8."""
import numpy as np
import pytest

import oracle
from oracle.cpu_baseline import CpuSupernodalCholesky
from gmrf_b200 import _lib, spde
from gmrf_b200.backend import B200Backend, _Handle
from gmrf_b200.introspect import Tables

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def _oracle_colcounts(Q, perm):
    """Exact column counts of L by the oracle's own symbolic pass (row subtrees; no numeric work)."""
    Q = Q.tocsc()
    n = Q.shape[0]
    parent, cc = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    nnz = oracle.lib().oracle_symbolic(n, Q.indptr.astype(np.int64), Q.indices.astype(np.int64),
                                       np.ascontiguousarray(perm, dtype=np.int64).ctypes.data, parent, cc)
    return cc, int(nnz)


def _backward_error(Q, x, b):
    return np.linalg.norm(Q @ x - b) / (np.linalg.norm(b) + abs(Q).sum(axis=1).max() * np.linalg.norm(x))


@pytest.mark.parametrize("cells,need_large_tile", [(40, False), (56, True)])
def test_3d_bench_paths_against_cpu_port(cells, need_large_tile):
    model = spde.MaternSPDE(*spde.mesh3d(cells), 0)
    perm = spde.geometric_nd_perm((cells + 1,) * 3, leaf=64, width=2)        # the ordering bench.py uses
    Q = model.precision(0.8, 0.6)
    n = Q.shape[0]
    be = B200Backend(Q, ordering=perm, device=0)
    info = be.info()
    # the bench's code paths are live at this size
    assert info["splitk_tasks"] >= 1, info
    if need_large_tile:
        assert info["large_tile_launches"] >= 1, info
    # symbolic parity: exact column counts from the oracle's independent pass under the same ordering
    cc, nnzl = _oracle_colcounts(Q, be.permutation())
    assert np.array_equal(be.colcounts(), cc) and info["nnz_l"] == nnzl
    # numeric parity against the CPU port on the same tables
    h = _Handle(n, Q.indptr, Q.indices, perm, _lib.ORDER_ND, device=-1)
    cpu = CpuSupernodalCholesky(Tables(h))
    cpu.refactorize(Q.data)
    assert cpu.status == 0 and be.status == 0
    assert abs(be.compute_logdet() - cpu.logdet) <= 1e-10 * abs(cpu.logdet)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    x = be.backend_solve(b)
    assert _backward_error(Q, x, b) <= 1e-10
    assert _rel(x, cpu.solve(b)[0]) <= 1e-8
    z = rng.standard_normal(n)
    assert _rel(be.backend_backward_solve(z), cpu.solve(z, half=True)[0]) <= 1e-8
    B70 = rng.standard_normal((n, 70))                                        # wide path: one 64-column block + a 6-column tail
    X70 = be.backend_solve(B70)
    assert _backward_error(Q, X70, B70) <= 1e-10
    for c in (0, 63, 64, 69):
        assert _rel(X70[:, c], be.backend_solve(B70[:, c])) <= 1e-10
    S70 = be.backend_backward_solve(B70)
    for c in (0, 69):
        assert _rel(S70[:, c], cpu.solve(B70[:, c], half=True)[0]) <= 1e-8
    # marginal variances: the port's supernodal Takahashi recursion
    var = be.get_selinv_diag()
    assert be.info()["fast_roots"] >= 1
    cpu.selinv()
    ref = cpu.selinv_diag()
    assert np.max(np.abs(var - ref) / ref) <= 1e-8
    # ... and independently of any recursion: unit-vector solves at sampled indices
    idx = rng.choice(n, 6, replace=False)
    E = np.zeros((n, 6)); E[idx, np.arange(6)] = 1.0
    assert np.allclose(var[idx], be.backend_solve(E)[idx, np.arange(6)], rtol=1e-8)
    # logdet(2Q) = logdet(Q) + n log 2
    ld = be.compute_logdet()
    Q2 = Q.copy(); Q2.data *= 2.0
    be.refactorize(Q2)
    assert abs(be.compute_logdet() - (ld + n * np.log(2.0))) <= 1e-10 * abs(ld)
    be.close(); h.close()


def _identities(Q, perm, rng, nsamp=5):
    """Size-independent properties standing in for the oracle at sizes where it takes minutes."""
    n = Q.shape[0]
    be = B200Backend(Q, ordering=perm, device=0)
    assert be.status == 0
    ld = be.compute_logdet()
    b = rng.standard_normal(n)
    x = be.backend_solve(b)
    assert _backward_error(Q, x, b) <= 1e-10
    var = be.get_selinv_diag()
    idx = rng.choice(n, nsamp, replace=False)
    E = np.zeros((n, nsamp)); E[idx, np.arange(nsamp)] = 1.0
    assert np.allclose(var[idx], be.backend_solve(E)[idx, np.arange(nsamp)], rtol=1e-8)
    z = rng.standard_normal(n)
    s = be.backend_backward_solve(z)                  # x = P' L^-T z  =>  x' Q x = z' z
    assert abs(s @ (Q @ s) - z @ z) <= 1e-9 * (z @ z)
    Q2 = Q.copy(); Q2.data *= 2.0
    be.refactorize(Q2)
    assert abs(be.compute_logdet() - (ld + n * np.log(2.0))) <= 1e-10 * abs(ld)
    info = be.info()
    be.close()
    return info


@pytest.mark.parametrize("cells", [500, 316])
def test_config2_and_3_sizes_identities(cells):
    """BASELINE configs 2 (251,001 dofs) and 3 (100,489 dofs): 2D Matern alpha = 3."""
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3)
    perm = spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3)
    info = _identities(Q, perm, np.random.default_rng(cells))
    assert info["front_launches"] >= 1 and info["chain_launches"] >= 1, info      # the latency paths are the ones checked


def test_config5_510k_identities():
    """BASELINE config 5 at 101^2 x 50 = 510,050 latent dofs: space-time advection-diffusion posterior."""
    coords, cells = spde.mesh2d(100)
    nt = 50
    model = spde.AdvectionDiffusionSSM(coords, cells, nt=nt)
    rng = np.random.default_rng(3)
    obs = rng.choice(model.ns, 500, replace=False)
    Q = model.posterior(obs, 1.0 / 0.05 ** 2)
    perm = spde.geometric_nd_perm((101, 101, nt), leaf=64, width=(5, 5, 1))
    _identities(Q, perm, rng, nsamp=3)
