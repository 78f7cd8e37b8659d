"""CPU tests (no GPU): symbolic analysis of the library vs the oracle, and a host replay of its schedule tables."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import replay
from gmrf_b200 import _lib, spde
from gmrf_b200.backend import _Handle, PinDenseColumns, ordering_permutation


def _handle(Q, perm=None, ordering=_lib.ORDER_ND):
    Q = sp.csc_matrix(Q)
    Q.sort_indices()
    return _Handle(Q.shape[0], Q.indptr, Q.indices, perm, ordering, device=-1)


CASES = {
    "grid_border": lambda: spde.grid_border_fixture(),
    "grid3d_6": lambda: spde.grid3d_fixture(6, 6, 6),
    "tridiag10": lambda: spde.tridiag_fixture(10),
    "rand20": lambda: spde.random_spd_fixture(20),
    "rand400": lambda: spde.random_spd_fixture(400, 0.02, 1),
    "matern2d_16": lambda: spde.MaternSPDE(*spde.mesh2d(16), 1).precision(1.0, 0.5),
    "matern3d_6": lambda: spde.MaternSPDE(*spde.mesh3d(6), 0).precision(1.0, 0.5),
    "diag": lambda: sp.identity(7, format="csc") * 2.0,
    "one": lambda: sp.csc_matrix(np.array([[3.0]])),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("ordering", [_lib.ORDER_ND, _lib.ORDER_AMD, _lib.ORDER_NATURAL])
def test_colcounts_exact_vs_oracle(name, ordering):
    Q = CASES[name]()
    h = _handle(Q, ordering=ordering)
    T = replay.Tables(h)
    F = oracle.OracleFactor(Q, T.perm)
    assert np.array_equal(T.colcount, F.colcount)          # bit-exact column counts under the same ordering
    assert np.array_equal(T.parent, F.parent)
    assert T.info["nnz_l"] == F.nnzL
    replay.check_structure(T)
    h.close()


@pytest.mark.parametrize("name", list(CASES))
def test_schedule_replay_matches_oracle(name):
    Q = CASES[name]()
    n = Q.shape[0]
    h = _handle(Q)
    T = replay.Tables(h)
    F = oracle.OracleFactor(Q, T.perm)
    Lx = replay.factor(T, Q.data)
    assert abs(replay.logdet(T, Lx) - F.logdet()) <= 1e-12 * max(1.0, abs(F.logdet()))
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    x = replay.solve(T, Lx, b)
    assert np.linalg.norm(x - F.solve(b)) <= 1e-10 * np.linalg.norm(x)
    z = rng.standard_normal(n)
    assert np.linalg.norm(replay.solve(T, Lx, z, half=True) - F.backward_solve(z)) <= 1e-10 * np.linalg.norm(z)
    Zx = replay.selinv(T, Lx)
    d = replay.selinv_diag(T, Zx)
    assert np.max(np.abs(d - F.selinv_diag()) / F.selinv_diag()) <= 1e-9
    h.close()


def test_user_permutation_and_errors():
    Q = spde.grid_border_fixture()
    n = Q.shape[0]
    perm = np.arange(n)[::-1].copy()
    h = _handle(Q, perm=perm)
    T = replay.Tables(h)
    # the final order is the user's order composed with an etree postorder: column counts are a permutation of the oracle's
    F = oracle.OracleFactor(Q, perm)
    assert sorted(T.colcount) == sorted(F.colcount)
    h.close()
    with pytest.raises(ValueError):
        _handle(Q, perm=np.zeros(n, dtype=np.int64))


def test_pin_dense_columns():
    Q = spde.grid_border_fixture()
    n = Q.shape[0]
    p = ordering_permutation(Q, PinDenseColumns("nd"))
    assert np.array_equal(np.sort(p), np.arange(n))
    assert p[-1] == n - 1                       # the dense border column goes last
    p2 = ordering_permutation(Q[: n - 1][:, : n - 1], PinDenseColumns("nd"))
    assert np.array_equal(np.sort(p2), np.arange(n - 1))


def test_numeric_calls_fail_without_device():
    Q = spde.tridiag_fixture(10)
    h = _handle(Q)
    nz = np.ascontiguousarray(Q.data)
    rc = h._L.gmrf_b200_refactorize(h._h, _lib.ptr(nz), nz.size)
    assert rc == -3                             # GMRF_B200_ERR_NO_DEVICE: no CPU fallback
    assert b"no CPU fallback" in h._L.gmrf_b200_last_error(h._h)
    h.close()


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "gmrf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gmrf_b200_[a-zA-Z_0-9]+)\s*\(", hdr)))
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert set(declared) == set(_lib.EXPORTS)


def test_threaded_host_copy_matches_numpy():
    """update_precision_values mirrors the caller's values into ws.Q.nzval; large arrays go through the library's threaded
    copy (a host utility: no device involved)."""
    rng = np.random.default_rng(4)
    for n in (0, 1, 1000, (1 << 20) + 12345, 3 * (1 << 20) + 7):
        src = rng.standard_normal(n)
        dst = np.full(n + 2, 7.0)
        assert _lib.lib().gmrf_b200_host_copy(_lib.ptr(dst[1:]) if n else None, _lib.ptr(src) if n else None, n) == 0
        assert np.array_equal(dst[1:n + 1], src) and dst[0] == 7.0 and dst[-1] == 7.0
    big = rng.standard_normal((1 << 21) + 3)
    out = np.zeros_like(big)
    _lib.host_copy(out, big)
    assert np.array_equal(out, big)
    _lib.host_copy(out[::2], big[::2] * 2.0)            # strided views fall back to numpy
    assert np.array_equal(out[::2], big[::2] * 2.0)
    assert _lib.lib().gmrf_b200_host_copy(None, None, -1) != 0


def test_amd_ordering_quality_and_speed():
    """Approximate minimum degree (supervariables, element absorption, mass elimination): a valid permutation whose
    fill stays within the usual AMD/ND band on a 2D mesh, in time near-linear in nnz(Q) (the exact-degree variant it
    replaced needed minutes at this size)."""
    import time
    Q = spde.MaternSPDE(*spde.mesh2d(160), 1).precision(1.0, 0.5)
    n = Q.shape[0]
    t = time.time()
    h_amd = _handle(Q, ordering=_lib.ORDER_AMD)
    dt = time.time() - t
    h_nd = _handle(Q, ordering=_lib.ORDER_ND)
    h_nat = _handle(Q, ordering=_lib.ORDER_NATURAL)
    p = h_amd.perm()
    assert np.array_equal(np.sort(p), np.arange(n))
    a, d, nat = h_amd.info()["nnz_l"], h_nd.info()["nnz_l"], h_nat.info()["nnz_l"]
    assert a <= 1.5 * d and a < 0.5 * nat, (a, d, nat)
    assert dt < 20.0, dt
    for h in (h_amd, h_nd, h_nat):
        h.close()


def test_amd_on_graphs_with_indistinguishable_and_isolated_nodes():
    """Cliques (every vertex indistinguishable -> one supervariable), stars (mass elimination) and isolated vertices."""
    blocks = [sp.csc_matrix(np.ones((6, 6))), sp.identity(3, format="csc")]
    star = sp.lil_matrix((9, 9))
    star[0, :] = 1.0
    star[:, 0] = 1.0
    blocks.append(star.tocsc())
    Q = sp.block_diag(blocks, format="csc") + 20.0 * sp.identity(18, format="csc")
    h = _handle(Q, ordering=_lib.ORDER_AMD)
    T = replay.Tables(h)
    assert np.array_equal(np.sort(T.perm), np.arange(18))
    F = oracle.OracleFactor(Q, T.perm)
    assert T.info["nnz_l"] == F.nnzL == 21 + 3 + 17   # no fill: clique 6*7/2, 3 singletons, star hub last (9 + 8)
    h.close()


@pytest.mark.parametrize("name", ["grid_border", "rand400", "matern2d_16", "matern3d_6", "diag", "one"])
def test_factor_export_pattern_is_the_square_root_layout(name):
    """`gmrf_b200_factor_pattern` (P'L as CSC, sparse_cho_sqrt of src/linear_maps/cholesky_sqrt.jl:6-21) on an
    analysis-only handle: sorted original row indices per pivot column, a superset of the exact pattern of L, and --
    filled from host-replayed panels through the panel layout -- a matrix R with R R' = Q."""
    Q = sp.csc_matrix(CASES[name]())
    n = Q.shape[0]
    h = _handle(Q)
    T = replay.Tables(h)
    cp, rv = h.factor_pattern()
    assert cp[0] == 0 and cp[-1] == rv.size and T.info["nnz_l"] <= rv.size <= T.info["nnz_l_stored"]
    Lx = replay.factor(T, Q.data)
    iperm = np.empty(n, dtype=np.int64)
    iperm[T.perm] = np.arange(n)
    vals = np.empty(rv.size)
    for s in range(T.nsuper):
        rows = T.rows(s)
        for lc in range(T.ns(s)):
            k = int(T.super_ptr[s]) + lc
            r = rv[cp[k]:cp[k + 1]]
            assert np.all(np.diff(r) > 0)                                  # sorted, no duplicates
            t = np.searchsorted(rows, iperm[r])
            assert np.array_equal(rows[t], iperm[r]) and np.all(t >= lc)   # inside the supernode's trapezoid
            assert r.size == T.nrow(s) - lc and T.perm[k] in r
            vals[cp[k]:cp[k + 1]] = Lx[int(T.panel_off[s]) + lc * int(T.panel_ld[s]) + t]
    R = sp.csc_matrix((vals, rv, cp), shape=(n, n))
    assert abs(R @ R.T - Q).max() <= 1e-12 * abs(Q).max()
    F = oracle.OracleFactor(Q, T.perm)                                     # exact pattern of L is contained in it
    Lo = sp.csc_matrix((F.Lx, F.Li, F.Lp), shape=(n, n))
    Ro = sp.csc_matrix(Lo[iperm, :])                                       # sparse(L)[invperm(p), :]
    pat = sp.csc_matrix((np.ones(rv.size), rv, cp), shape=(n, n))
    assert (Ro != 0).multiply(pat).nnz == (Ro != 0).nnz
    assert abs(R - Ro).max() <= 1e-10 * abs(Ro).max()
    h.close()


def _positions(h, B, base=0):
    B = sp.csc_matrix(B)
    B.sort_indices()
    cp = (B.indptr.astype(np.int64) + base)
    rv = (B.indices.astype(np.int64) + base)
    pos = np.empty(rv.size, dtype=np.int64)
    rc = h._L.gmrf_b200_pattern_positions(h._h, B.shape[1], _lib.ptr(cp), _lib.ptr(rv), base, _lib.ptr(pos))
    return rc, pos


def test_pattern_positions_lookup():
    """The host lookup behind selinv_extract / selinv_dot (OpenMP over columns): panel offsets of a caller pattern,
    against offsets derived independently from the exported tables; -1 outside the stored pattern; both index bases;
    symmetric in (i, j); usage errors."""
    Q = sp.csc_matrix(CASES["matern3d_6"]())
    n = Q.shape[0]
    h = _handle(Q)
    T = replay.Tables(h)
    iperm = np.empty(n, dtype=np.int64)
    iperm[T.perm] = np.arange(n)
    col2super = np.repeat(np.arange(T.nsuper), np.diff(T.super_ptr))
    rng = np.random.default_rng(0)
    B = sp.csc_matrix(Q + sp.random(n, n, density=0.02, random_state=rng, format="csc"))
    B.sort_indices()
    rc, pos = _positions(h, B)
    assert rc == 0
    C = B.tocoo()
    order = np.lexsort((C.row, C.col))                                  # CSC order
    rows, cols = C.row[order], C.col[order]
    want = np.empty(rows.size, dtype=np.int64)
    for k, (i, j) in enumerate(zip(rows, cols)):
        a, b = iperm[i], iperm[j]
        c, r = min(a, b), max(a, b)
        s = col2super[c]
        rs = T.rows(s)
        t = np.searchsorted(rs, r)
        want[k] = (T.panel_off[s] + (c - T.super_ptr[s]) * T.panel_ld[s] + t) if t < rs.size and rs[t] == r else -1
    assert np.array_equal(pos, want)
    assert (pos < 0).any() and (pos >= 0).any()
    rc1, pos1 = _positions(h, B, base=1)
    assert rc1 == 0 and np.array_equal(pos1, pos)
    rcT, posT = _positions(h, B.T)                                     # Sigma is symmetric: (j, i) reads the same entry
    CT = sp.csc_matrix(B.T).tocoo()
    oT = np.lexsort((CT.row, CT.col))
    lut = dict(zip(zip(rows.tolist(), cols.tolist()), pos.tolist()))
    assert rcT == 0 and all(lut[(j, i)] == p for i, j, p in zip(CT.row[oT].tolist(), CT.col[oT].tolist(), posT.tolist()))
    # entries of Q itself are the scatter map (upper triangle)
    rcq, posq = _positions(h, Q)
    assert rcq == 0 and np.array_equal(posq[T.q_src], T.q_dst)
    # errors: wrong size, row index out of range
    assert _positions(h, sp.identity(n + 1, format="csc"))[0] == -1
    bad = sp.csc_matrix(Q)
    cp, rv = bad.indptr.astype(np.int64), bad.indices.astype(np.int64).copy()
    rv[5] = n + 3
    out = np.empty(rv.size, dtype=np.int64)
    assert h._L.gmrf_b200_pattern_positions(h._h, n, _lib.ptr(cp), _lib.ptr(rv), 0, _lib.ptr(out)) == -1
    assert b"row index out of range" in h._L.gmrf_b200_last_error(h._h)
    h.close()


def test_pattern_positions_remembers_the_last_pattern():
    """A repeated pattern (gradient loops, linear-predictor marginals) skips the lookup: content-compared, one entry,
    replaced on any difference (pattern, index base); results identical either way."""
    Q = sp.csc_matrix(CASES["matern2d_16"]())
    n = Q.shape[0]
    h = _handle(Q)
    hits = lambda: h.info()["pattern_cache_hits"]
    B = sp.csc_matrix(Q + sp.random(n, n, density=0.05, random_state=np.random.default_rng(1), format="csc"))
    rc, p0 = _positions(h, Q)
    assert rc == 0 and hits() == 0
    rc, p1 = _positions(h, Q)
    assert rc == 0 and hits() == 1 and np.array_equal(p0, p1)
    rc, pb = _positions(h, B)                               # different pattern: miss, cache replaced
    assert rc == 0 and hits() == 1
    rc, pb2 = _positions(h, B)
    assert rc == 0 and hits() == 2 and np.array_equal(pb, pb2)
    rc, p2 = _positions(h, Q)                               # back to the first: miss again, same answer
    assert rc == 0 and hits() == 2 and np.array_equal(p2, p0)
    rc, p3 = _positions(h, Q, base=1)                       # same pattern, other index base: miss, same answer
    assert rc == 0 and hits() == 2 and np.array_equal(p3, p0)
    C = Q.copy()                                            # same sizes, one row index changed
    C.indices = C.indices.copy()
    j = int(np.flatnonzero(np.diff(C.indptr) >= 2)[0])
    C.indices[C.indptr[j]], C.indices[C.indptr[j] + 1] = C.indices[C.indptr[j] + 1], C.indices[C.indptr[j]]
    cp, rv = C.indptr.astype(np.int64), C.indices.astype(np.int64)
    out = np.empty(rv.size, dtype=np.int64)
    assert h._L.gmrf_b200_pattern_positions(h._h, n, _lib.ptr(cp), _lib.ptr(rv), 0, _lib.ptr(out)) == 0
    assert hits() == 2 and out[cp[j]] == p0[cp[j] + 1] and out[cp[j] + 1] == p0[cp[j]]
    # a failed lookup must not leave a stale entry behind
    rv_bad = rv.copy()
    rv_bad[0] = -7
    assert h._L.gmrf_b200_pattern_positions(h._h, n, _lib.ptr(cp), _lib.ptr(rv_bad), 0, _lib.ptr(out)) == -1
    assert h._L.gmrf_b200_pattern_positions(h._h, n, _lib.ptr(cp), _lib.ptr(rv_bad), 0, _lib.ptr(out)) == -1
    rc, p4 = _positions(h, Q)
    assert rc == 0 and np.array_equal(p4, p0)
    empty = sp.csc_matrix((n, n))
    assert _positions(h, empty)[0] == 0 and _positions(h, empty)[0] == 0
    h.close()


def test_create_rejects_patterns_that_break_the_csc_invariants():
    """SparseMatrixCSC invariants are checked up front (ArgumentError-like -1 with a message), never read past."""
    def rc_msg(n, cp, rv, perm=None, ordering=_lib.ORDER_ND):
        try:
            _Handle(n, np.asarray(cp, dtype=np.int64), np.asarray(rv, dtype=np.int64), perm, ordering, device=-1).close()
            return None
        except ValueError as e:
            return str(e)
    assert "strictly increasing" in rc_msg(2, [0, 3, 5], [0, 0, 1, 0, 1])            # duplicate entry
    assert "strictly increasing" in rc_msg(3, [0, 2, 5, 7], [1, 0, 2, 0, 1, 2, 1])   # unsorted rows
    assert "non-decreasing" in rc_msg(2, [0, 3, 2], [0, 1, 1])                       # colptr goes backwards
    assert "row index out of range" in rc_msg(2, [0, 2, 4], [0, 5, 0, 1])
    assert "not a permutation" in rc_msg(2, [0, 2, 4], [0, 1, 0, 1], perm=np.array([0, 0], dtype=np.int64))
    assert "unknown ordering" in rc_msg(2, [0, 2, 4], [0, 1, 0, 1], ordering=7)
    assert rc_msg(0, [0], []) is None                                                # the empty matrix is fine
    assert rc_msg(2, [0, 1, 2], [0, 1]) is None


def test_symbolic_analysis_fuzz_against_the_oracle():
    """Random sparse SPD patterns (disconnected pieces, dense rows, tiny n) x every ordering incl. random user
    permutations: exact column counts and etree equal the oracle's, the tables pass the structural invariants, and a
    host replay of the schedule reproduces the oracle's log-determinant."""
    rng = np.random.default_rng(7)
    for it in range(60):
        n = int(rng.integers(1, 120))
        dens = float(rng.choice([0.003, 0.02, 0.1, 0.4]))
        A = sp.random(n, n, dens, random_state=int(rng.integers(1 << 30)), format="csc")
        A = sp.csc_matrix(A + A.T + sp.identity(n) * n)
        A.sort_indices()
        for ordering in (_lib.ORDER_NATURAL, _lib.ORDER_ND, _lib.ORDER_AMD):
            perm = rng.permutation(n).astype(np.int64) if (it % 3 == 0 and ordering == _lib.ORDER_NATURAL) else None
            h = _handle(A, perm=perm, ordering=ordering)
            T = replay.Tables(h)
            F = oracle.OracleFactor(A, T.perm)
            assert np.array_equal(T.colcount, F.colcount) and np.array_equal(T.parent, F.parent)
            replay.check_structure(T)
            if it % 6 == 0:
                Lx = replay.factor(T, A.data)
                assert abs(replay.logdet(T, Lx) - F.logdet()) <= 1e-10 * max(1.0, abs(F.logdet()))
            h.close()


@pytest.mark.parametrize("name", ["grid_border", "rand400", "matern2d_16", "matern3d_6", "diag", "one"])
@pytest.mark.parametrize("ordering", [_lib.ORDER_ND, _lib.ORDER_AMD])
def test_analysis_round_trip(name, ordering):
    """export -> create_from_analysis: every table identical (checked member by member inside the library and through
    the introspection calls), the restored handle replays to the oracle's log-determinant, export is idempotent."""
    Q = sp.csc_matrix(CASES[name]())
    Q.sort_indices()
    n = Q.shape[0]
    cp, rv = Q.indptr.astype(np.int64), Q.indices.astype(np.int64)
    h = _Handle(n, cp, rv, None, ordering, device=-1)
    blob = h.export_analysis()
    h2 = _Handle(n, cp, rv, None, _lib.ORDER_NATURAL, device=-1, analysis=blob)      # ordering argument is ignored
    assert h._L.gmrf_b200_analysis_equal(h._h, h2._h) == 1
    assert h2.export_analysis() == blob
    T, T2 = replay.Tables(h), replay.Tables(h2)
    for a in ("perm", "colcount", "parent", "super_ptr", "sparent", "level", "row_ptr", "row_idx", "rel_idx", "panel_off",
              "panel_ld", "upd_off", "upd_ld", "q_src", "q_dst"):
        assert np.array_equal(getattr(T, a), getattr(T2, a)), a
    assert T.info == T2.info
    replay.check_structure(T2)
    F = oracle.OracleFactor(Q, T2.perm)
    assert abs(replay.logdet(T2, replay.factor(T2, Q.data)) - F.logdet()) <= 1e-12 * max(1.0, abs(F.logdet()))
    p1, p2 = h.factor_pattern(), h2.factor_pattern()
    assert np.array_equal(p1[0], p2[0]) and np.array_equal(p1[1], p2[1])
    h.close()
    h2.close()


def test_analysis_blob_is_tied_to_its_pattern_and_validated():
    Q = sp.csc_matrix(CASES["matern2d_16"]())
    Q.sort_indices()
    n = Q.shape[0]
    cp, rv = Q.indptr.astype(np.int64), Q.indices.astype(np.int64)
    h = _handle(Q)
    blob = h.export_analysis()
    other = sp.csc_matrix(CASES["grid_border"]())
    with pytest.raises(ValueError, match="different matrix"):
        _Handle(other.shape[0], other.indptr.astype(np.int64), other.indices.astype(np.int64), None, 0, device=-1, analysis=blob)
    # same size and nnz, two row indices of one column changed: the pattern hash catches it
    rv2 = rv.copy()
    j = int(np.flatnonzero(np.diff(cp) >= 3)[0])
    free = np.setdiff1d(np.arange(n), rv[cp[j]:cp[j + 1]])
    col = np.sort(np.concatenate([rv[cp[j]:cp[j + 1] - 1], free[-1:]]))
    rv2[cp[j]:cp[j + 1]] = col
    with pytest.raises(ValueError, match="different sparsity pattern"):
        _Handle(n, cp, rv2, None, 0, device=-1, analysis=blob)
    for bad, msg in ((blob[:-9], "truncated|malformed"), (blob + b"x", "trailing"), (b"not a blob at all!", "not an analysis blob"),
                     (blob[:8] + bytes(8) + blob[16:], "different sparsity pattern")):
        with pytest.raises(ValueError, match=msg):
            _Handle(n, cp, rv, None, 0, device=-1, analysis=bad)
    # numeric calls on the restored analysis-only handle still refuse to run without a device
    h2 = _Handle(n, cp, rv, None, 0, device=-1, analysis=blob)
    nz = np.ascontiguousarray(Q.data)
    assert h2._L.gmrf_b200_refactorize(h2._h, _lib.ptr(nz), nz.size) == -3
    h.close()
    h2.close()


def test_schedule_replay_at_10k_dofs_against_superlu():
    """The library's tables for a 10,201-dof 2D Matern precision (441 supernodes, 11 levels, relaxed amalgamation, pool
    reuse across levels) replayed on the host against an independent factorization (SciPy SuperLU): log-determinant,
    solve and selected-inversion diagonal."""
    import scipy.sparse.linalg as spl
    Q = sp.csc_matrix(spde.MaternSPDE(*spde.mesh2d(100), 1).precision(1.0, 0.3))
    Q.sort_indices()
    n = Q.shape[0]
    h = _handle(Q)
    T = replay.Tables(h)
    assert T.nsuper > 300 and T.info["nlevels"] >= 8
    Lx = replay.factor(T, Q.data)
    lu = spl.splu(Q, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0, options=dict(SymmetricMode=True))
    ld_ref = float(np.sum(np.log(np.abs(lu.U.diagonal()))))
    assert abs(replay.logdet(T, Lx) - ld_ref) <= 1e-10 * abs(ld_ref)
    b = np.random.default_rng(0).standard_normal(n)
    x = replay.solve(T, Lx, b)
    assert np.linalg.norm(Q @ x - b) <= 1e-10 * (np.linalg.norm(b) + abs(Q).sum(axis=1).max() * np.linalg.norm(x))
    d = replay.selinv_diag(T, replay.selinv(T, Lx))
    idx = np.array([0, n // 3, n // 2, n - 1])
    E = np.zeros((n, idx.size))
    E[idx, np.arange(idx.size)] = 1.0
    ref = lu.solve(E)[idx, np.arange(idx.size)]
    assert np.max(np.abs(d[idx] - ref) / ref) <= 1e-8
    h.close()


def test_analysis_tables_do_not_depend_on_the_thread_count():
    """Above 10^5 columns the analysis runs its counting sorts, row structures, scatter map and position lookups on all
    cores; every table (and the exported analysis stream) must be identical to the single-threaded result."""
    import hashlib
    import os
    import subprocess
    import sys
    script = (
        "import sys, hashlib, numpy as np\n"
        "sys.path[:0] = [%r, %r]\n"
        "from gmrf_b200 import _lib, spde\n"
        "from gmrf_b200.backend import _Handle\n"
        "Q = spde.MaternSPDE(*spde.mesh2d(330), 1).precision(1.0, 0.3); Q.sort_indices(); n = Q.shape[0]\n"
        "cp, rv = Q.indptr.astype(np.int64), Q.indices.astype(np.int64)\n"
        "m = hashlib.sha256()\n"
        "for o in (_lib.ORDER_AMD, _lib.ORDER_NATURAL):\n"
        "    h = _Handle(n, cp, rv, None, o, device=-1); m.update(h.export_analysis()); h.close()\n"
        "h = _Handle(n, cp, rv, spde.geometric_nd_perm((331, 331), leaf=64, width=3), 1, device=-1)\n"
        "pos = np.empty(rv.size, dtype=np.int64)\n"
        "assert h._L.gmrf_b200_pattern_positions(h._h, n, _lib.ptr(cp), _lib.ptr(rv), 0, _lib.ptr(pos)) == 0\n"
        "m.update(h.export_analysis()); m.update(pos.tobytes()); print(n, m.hexdigest())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
         os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaussianmarkovrandomfields.jl_b200"))
    outs = []
    for threads in ("1", "4"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and outs[0].startswith("109561 ")


def test_pool_can_share_one_analysis():
    """`WorkspacePool(..., share_analysis=True)`: the first slot analyses, the others are created from its exported
    analysis (analysis-only handles here: device = -1)."""
    from gmrf_b200.workspace import WorkspacePool
    Q = sp.csc_matrix(CASES["matern2d_16"]())
    pool = WorkspacePool(Q, size=3, devices=(-1,), share_analysis=True, factorize=False)
    hs = [ws.backend._hd for ws in pool.workspaces]
    assert all(hs[0]._L.gmrf_b200_analysis_equal(hs[0]._h, h._h) == 1 for h in hs[1:])
    assert all(np.array_equal(ws.backend.permutation(), pool.workspaces[0].backend.permutation()) for ws in pool.workspaces)
    plain = WorkspacePool(Q, size=2, devices=(-1,), factorize=False)
    assert hs[0]._L.gmrf_b200_analysis_equal(hs[0]._h, plain.workspaces[1].backend._hd._h) == 1    # same result either way


def test_one_based_create_gives_the_same_analysis():
    """index_base = 1 (the Julia glue's arrays) and index_base = 0 describe the same matrix: identical tables, identical
    (0-based) permutation, identical exported patterns -- host-side, no GPU."""
    from gmrf_b200.backend import index_base
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    A = sp.random(60, 60, density=0.08, random_state=rng, format="csc")
    Q = sp.csc_matrix(A + A.T + 10 * sp.identity(60)); Q.sort_indices()
    user = rng.permutation(60)
    for ordering in (None, user):
        hs = []
        for base in (0, 1):
            with index_base(base):
                h = _Handle(60, Q.indptr, Q.indices, ordering, _lib.ORDER_AMD, device=-1)
                hs.append((h, h.perm(), h.factor_pattern()))
        (h0, p0, f0), (h1, p1, f1) = hs
        assert h0._L.gmrf_b200_analysis_equal(h0._h, h1._h) == 1
        assert np.array_equal(p0, p1) and np.array_equal(f0[0], f1[0]) and np.array_equal(f0[1], f1[1])
        if ordering is not None:
            assert np.array_equal(np.sort(p0), np.arange(60))
        h0.close(); h1.close()


def test_pivot_tile_recurrence_restated_on_the_host():
    """Host restatement of the two device routines under every panel factorization (csrc/kernels.cuh): `chol4x4_lower`
    factors a 4 x 4 pivot tile as two 2 x 2 blocks whose second pivot comes from the block's determinant, and
    `rsqrt_inline` refines a ~20-bit hardware seed with one third-order step. The claims made for them -- backward stable
    like the column recurrence, 1 ulp-ish -- are checked here in plain numpy (the GPU suite checks the device code itself)."""
    rng = np.random.default_rng(12)

    def rsq(x):                       # seed rounded to 20 bits, then y (1 + e/2 + 3 e^2/8), e = 1 - x y^2
        m, ex = np.frexp(1.0 / np.sqrt(x))
        y = np.ldexp(np.round(m * 2.0 ** 20) / 2.0 ** 20, ex)
        e = np.float64(1.0 - np.longdouble(x) * y * y)            # (fma on the device)
        return y * (e * (0.375 * e + 0.5)) + y

    xs = 10.0 ** rng.uniform(-150, 150, 20000)
    exact = 1.0 / np.sqrt(xs.astype(np.longdouble))
    assert np.max(np.abs((rsq(xs).astype(np.longdouble) - exact) / exact)) <= 2.5e-16

    worst = 0.0
    for _ in range(3000):
        M = rng.standard_normal((4, 6))
        A = M @ M.T + 10.0 ** rng.uniform(-9, 0) * np.eye(4)
        a = np.tril(A)
        det1 = a[0, 0] * a[1, 1] - a[1, 0] * a[1, 0]
        r0, q1 = rsq(a[0, 0]), rsq(det1)
        l00, l10 = a[0, 0] * r0, a[1, 0] * r0
        ri1 = l00 * q1
        l20, l30 = a[2, 0] * r0, a[3, 0] * r0
        l21, l31 = (a[2, 1] - l20 * l10) * ri1, (a[3, 1] - l30 * l10) * ri1
        s22 = a[2, 2] - l20 * l20 - l21 * l21
        s32 = a[3, 2] - l30 * l20 - l31 * l21
        s33 = a[3, 3] - l30 * l30 - l31 * l31
        det2 = s22 * s33 - s32 * s32
        r2, q3 = rsq(s22), rsq(det2)
        L = np.array([[l00, 0, 0, 0], [l10, det1 * q1 * r0, 0, 0], [l20, l21, s22 * r2, 0], [l30, l31, s32 * r2, det2 * q3 * r2]])
        ri = np.array([r0, ri1, r2, s22 * r2 * q3])
        worst = max(worst, np.max(np.abs(L @ L.T - A)) / np.max(np.abs(A)))
        assert np.allclose(ri * np.diag(L), 1.0, rtol=1e-13)        # the published reciprocal pivots
        assert min(a[0, 0], det1, s22, det2) > 0                      # the four pivot-sign indicators
    assert worst <= 4e-15                                             # backward error of the tile factorization
    # an indefinite tile is flagged at the first failing column by the sign of the matching indicator
    A = np.diag([2.0, 3.0, 4.0, 5.0]); A[1, 0] = A[0, 1] = 3.0        # a00 a11 - a10^2 = -3 < 0 -> column 2 (1-based)
    assert A[0, 0] > 0 and A[0, 0] * A[1, 1] - A[1, 0] ** 2 < 0
