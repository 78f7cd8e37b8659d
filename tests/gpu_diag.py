"""Staged GPU diagnosis (run on the GPU box): compares every phase with the host replay / oracle and reports the
first supernode level that diverges. Not a pytest file."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
import oracle, replay
from gmrf_b200 import spde, _lib
from gmrf_b200._lib import ptr
from gmrf_b200.backend import B200Backend


def diag(name, Q, naive=0, graph=0):
    _lib.set_option("naive_kernels", naive)
    _lib.set_option("use_graph", graph)
    n = Q.shape[0]
    t0 = time.time()
    b = B200Backend(Q, device=0)
    T = replay.Tables(b._hd)
    info = b.info()
    Lx_ref = replay.factor(T, Q.data)
    Lx = np.empty(info["nnz_l_stored"])
    rc = b._L.gmrf_b200_get_factor_panels(b._hd._h, ptr(Lx), Lx.size)
    assert rc == 0, b._L.gmrf_b200_last_error(b._hd._h)
    worst = {}
    for s in range(T.nsuper):
        P, R = T.panel(Lx, s), T.panel(Lx_ref, s)
        ns = T.ns(s)
        mask = np.ones_like(P, dtype=bool)
        mask[:ns, :ns] = np.tril(np.ones((ns, ns), dtype=bool))
        err = np.abs(np.where(mask, P - R, 0.0)).max() / max(np.abs(R).max(), 1e-300)
        lv = int(T.level[s])
        worst[lv] = max(worst.get(lv, 0.0), err if np.isfinite(err) else np.inf)
    bad = [lv for lv in sorted(worst) if not worst[lv] < 1e-9]
    print(f"[{name} naive={naive} graph={graph}] n={n} nsuper={T.nsuper} levels={info['nlevels']} status={b.status} "
          f"factor max rel err per level: " + " ".join(f"{lv}:{worst[lv]:.1e}" for lv in sorted(worst)))
    if bad:
        print("   FIRST BAD LEVEL", bad[0])
    F = oracle.OracleFactor(Q, T.perm)
    ld = b.compute_logdet()
    print(f"   logdet gpu={ld:.15g} oracle={F.logdet():.15g} rel={abs(ld - F.logdet()) / abs(F.logdet()):.2e}")
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal(n)
    x = b.backend_solve(rhs); xr = F.solve(rhs)
    print(f"   solve rel err {np.linalg.norm(x - xr) / np.linalg.norm(xr):.2e}  residual {np.linalg.norm(Q @ x - rhs) / np.linalg.norm(rhs):.2e} (oracle residual {np.linalg.norm(Q @ xr - rhs) / np.linalg.norm(rhs):.2e})")
    R3 = rng.standard_normal((n, 11))
    X3 = b.backend_solve(R3); X3r = F.solve(R3)
    print(f"   solve 11 rhs rel err {np.linalg.norm(X3 - X3r) / np.linalg.norm(X3r):.2e}")
    z = rng.standard_normal(n)
    s = b.backend_backward_solve(z); sr = F.backward_solve(z)
    print(f"   Lt-solve rel err {np.linalg.norm(s - sr) / np.linalg.norm(sr):.2e}")
    d = b.get_selinv_diag(); dr = F.selinv_diag()
    print(f"   selinv diag max rel err {np.max(np.abs(d - dr) / dr):.2e}")
    Z = b.get_selinv(); Zr = F.selinv()
    # compare on the oracle's (exact) pattern
    Zr = Zr.tocoo()
    got = np.asarray(Z[Zr.row, Zr.col]).ravel()
    print(f"   selinv full max abs err {np.abs(got - Zr.data).max():.2e} (scale {np.abs(Zr.data).max():.2e}) nnzZ={Z.nnz}")
    print(f"   timings {b.timings()}  wall {time.time() - t0:.2f}s")
    b.close()


if __name__ == "__main__":
    cases = [
        ("tridiag10", spde.tridiag_fixture(10)),
        ("grid_border", spde.grid_border_fixture()),
        ("rand400", spde.random_spd_fixture(400, 0.02, 1)),
        ("dense600", spde.random_spd_fixture(600, 0.3, 3)),
        ("matern3d_16", spde.MaternSPDE(*spde.mesh3d(16), 0).precision(1.0, 0.5)),
        ("matern2d_32", spde.MaternSPDE(*spde.mesh2d(32), 1).precision(1.0, 0.5)),
        ("matern3d_10", spde.MaternSPDE(*spde.mesh3d(10), 0).precision(1.0, 0.5)),
    ]
    for naive, graph in ((0, 0), (0, 1)):
        for name, Q in cases:
            try:
                diag(name, Q, naive, graph)
            except Exception as e:
                print(f"[{name} naive={naive} graph={graph}] EXCEPTION {type(e).__name__}: {e}")
