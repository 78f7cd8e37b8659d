"""CPU-baseline timings of the BASELINE.json 2D configs with the supernodal CPU port in oracle/ (the stand-in for the
reference's CHOLMOD path; not a pytest file, no GPU needed):   python tests/cpu_configs.py [1] [2] [3]
One JSON line per config: numeric factorization + logdet, one solve, one half solve, selected inversion, on all host
threads, same inputs and the same (geometric nested dissection) ordering as tests/gpu_configs.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import _lib, spde  # noqa: E402
from gmrf_b200.backend import _Handle  # noqa: E402
from gmrf_b200.introspect import Tables  # noqa: E402
from oracle.cpu_baseline import CpuSupernodalCholesky  # noqa: E402

CELLS = {1: 224, 2: 500, 3: 316}
for c in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 3]:
    cells = CELLS[c]
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3)
    n = Q.shape[0]
    t0 = time.perf_counter()
    h = _Handle(n, Q.indptr.astype(np.int64), Q.indices.astype(np.int64),
                spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), _lib.ORDER_ND, device=-1)
    T = Tables(h)
    t_analysis = time.perf_counter() - t0
    cpu = CpuSupernodalCholesky(T)
    cpu.refactorize(Q.data)
    t_factor = min(cpu.refactorize(Q.data) for _ in range(3))
    rhs = np.random.default_rng(0).standard_normal(n)
    x, _ = cpu.solve(rhs)
    t_solve = min(cpu.solve(rhs)[1] for _ in range(7))
    t_half = min(cpu.solve(rhs, half=True)[1] for _ in range(7))
    t_selinv = min(cpu.selinv() for _ in range(2))
    flops = float(T.info["flops_chol"])
    print(json.dumps({"config": c, "n": n, "threads": cpu.threads, "analysis_s": round(t_analysis, 2),
                      "factor_logdet_ms": round(1e3 * t_factor, 2), "factor_gflops": round(flops / t_factor / 1e9, 1),
                      "solve_ms": round(1e3 * t_solve, 2), "half_solve_ms": round(1e3 * t_half, 2),
                      "selinv_ms": round(1e3 * t_selinv, 1),
                      "residual": float(np.linalg.norm(Q @ x - rhs) / np.linalg.norm(rhs)), "logdet": cpu.logdet}), flush=True)
    h.close()
    if "--superlu" in sys.argv:
        # independent third-party data point: SciPy SuperLU, single thread, LU not Cholesky
        import scipy.sparse.linalg as spl
        t0 = time.perf_counter()
        lu = spl.splu(Q, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0, options=dict(SymmetricMode=True))
        t_lu = time.perf_counter() - t0
        t0 = time.perf_counter()
        xs = lu.solve(rhs)
        t_s = time.perf_counter() - t0
        print(json.dumps({"config": c, "superlu_factor_ms": round(1e3 * t_lu, 1), "superlu_solve_ms": round(1e3 * t_s, 2),
                          "superlu_logdet": float(np.sum(np.log(np.abs(lu.U.diagonal())))),
                          "solution_rel_diff_vs_port": float(np.linalg.norm(xs - x) / np.linalg.norm(x))}), flush=True)
    if c == 2 and "--newton" in sys.argv:
        # the config's actual workload: the Poisson Newton loop through the host mirror, CPU port as the backend
        from cpu_port_backend import CpuPortBackend
        from gmrf_b200.workspace_gmrf import PoissonLikelihood, WorkspaceGMRF, gaussian_approximation
        coords = spde.mesh2d(cells)[0]
        lam = np.exp(0.5 + 0.5 * np.sin(2 * np.pi * coords[:, 0]) * np.cos(2 * np.pi * coords[:, 1]))
        lik = PoissonLikelihood(np.random.default_rng(1).poisson(lam))
        prior = WorkspaceGMRF(np.zeros(n), Q, backend_type=CpuPortBackend,
                              ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3))
        stats = {}
        t0 = time.perf_counter()
        post = gaussian_approximation(prior, lik, stats=stats)
        wall = time.perf_counter() - t0
        g = Q @ post.mean() - lik.loggrad(post.mean())
        print(json.dumps({"config": 2, "workload": "Poisson Newton loop", "n": n, "newton_wall_s": round(wall, 2), **stats,
                          "grad_inf": float(np.max(np.abs(g)))}), flush=True)
