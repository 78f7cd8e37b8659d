"""BASELINE.json configs 1, 2, 3, 5 end to end on one B200 (config 4 is bench.py). Not a pytest file; prints one JSON line
per config with timings (CUDA-event times from the library where they exist, wall clock otherwise) and the size-independent
checks that stand in for parity at these sizes.   python tests/gpu_configs.py [1] [2] [3] [5] [--small]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gmrf_b200.workspace import GMRFWorkspace  # noqa: E402
from gmrf_b200.workspace_gmrf import PoissonLikelihood, WorkspaceGMRF, gaussian_approximation  # noqa: E402

small = "--small" in sys.argv
which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 3, 5]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def _peaks():
    """Roofline denominators: measured HBM copy rate (MEASURED_PEAKS.json, else the round-1 figure) and the measured
    cuBLAS DGEMM rate (profiles/r01_fp64_probe.json)."""
    hbm, fp64 = 6541.5, 36.086
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    return hbm, fp64


def roofline(info, factor_ms=None, solve_ms=None, half_solve_ms=None, selinv_ms=None, nrhs=1):
    """Achieved rates against the roofline that bounds each phase: factorization / selected inversion in
    algorithmic FP64 TFLOP/s (sum cc_j^2, and 2x that), few-RHS solves in GB/s of factor panels streamed."""
    hbm, fp64 = _peaks()
    out = {}
    flops, lbytes = float(info["flops_chol"]), 8.0 * float(info["nnz_l_stored"])
    if factor_ms:
        out["factor_tflops"] = round(flops / (factor_ms * 1e-3) / 1e12, 3)
        out["factor_frac_of_dgemm"] = round(out["factor_tflops"] / fp64, 4)
    if selinv_ms:
        out["selinv_tflops_equiv"] = round(2.0 * flops / (selinv_ms * 1e-3) / 1e12, 3)
        out["selinv_frac_of_dgemm"] = round(out["selinv_tflops_equiv"] / fp64, 4)
    if solve_ms:
        out["solve_GBs"] = round(2.0 * lbytes / (solve_ms * 1e-3) / 1e9, 1)
        out["solve_frac_of_hbm"] = round(out["solve_GBs"] / hbm, 4)
    if half_solve_ms:
        out["half_solve_GBs"] = round(lbytes / (half_solve_ms * 1e-3) / 1e9, 1)
        out["half_solve_frac_of_hbm"] = round(out["half_solve_GBs"] / hbm, 4)
    return out


def config1():
    cells = 64 if small else 224
    coords, tri = spde.mesh2d(cells)
    model = spde.MaternSPDE(coords, tri, 1)
    Q = model.precision(1.0, 0.3)
    n = Q.shape[0]
    t0 = time.perf_counter()
    ws = GMRFWorkspace(Q, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0)
    setup = time.perf_counter() - t0
    be = ws.backend
    rng = np.random.default_rng(0)
    times = {}
    for _ in range(3):
        be.refactorize(Q); times["factor_logdet_ms"] = be.timings()["factor_ms"]
    ld = ws.logdet()
    b = rng.standard_normal(n)
    for _ in range(2):
        x = ws.workspace_solve(b); times["mean_solve_ms"] = be.timings()["solve_ms"]
    z = rng.standard_normal(n)
    for _ in range(2):
        s = ws.backward_solve(z); times["rand_ms"] = be.timings()["solve_ms"]
    be.refactorize(Q); be.selinv_compute(); times["selinv_ms"] = be.timings()["selinv_ms"]
    std = np.sqrt(ws.selinv_diag())
    idx = rng.choice(n, 4, replace=False)
    E = np.zeros((n, 4)); E[idx, np.arange(4)] = 1.0
    ref = ws.workspace_solve(E)[idx, np.arange(4)]
    emit(config=1, n=n, nnz_q=int(Q.nnz), setup_s=round(setup, 2), **{k: round(v, 3) for k, v in times.items()},
         logdet=ld, residual=float(np.linalg.norm(Q @ x - b) / np.linalg.norm(b)),
         std_vs_unit_solves=float(np.max(np.abs(std[idx] ** 2 - ref) / ref)), info=be.info()["graph_nodes"],
         roofline=roofline(be.info(), times["factor_logdet_ms"], times["mean_solve_ms"], times["rand_ms"], times["selinv_ms"]))


def config2():
    cells = 100 if small else 500
    coords, tri = spde.mesh2d(cells)
    model = spde.MaternSPDE(coords, tri, 1)
    Q = model.precision(1.0, 0.3)
    n = Q.shape[0]
    lam = np.exp(0.5 + 0.5 * np.sin(2 * np.pi * coords[:, 0]) * np.cos(2 * np.pi * coords[:, 1]))
    lik = PoissonLikelihood(np.random.default_rng(1).poisson(lam))
    t0 = time.perf_counter()
    prior = WorkspaceGMRF(np.zeros(n), Q, backend_type=B200Backend, device=0,
                          ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3))
    setup = time.perf_counter() - t0
    stats = {}
    t0 = time.perf_counter()
    post = gaussian_approximation(prior, lik, stats=stats)
    wall = time.perf_counter() - t0
    be = prior.workspace.backend
    g = Q @ (post.mean() - 0.0) - lik.loggrad(post.mean())
    t1 = time.perf_counter(); sd = post.std(); t_std = time.perf_counter() - t1
    emit(config=2, n=n, nnz_q=int(Q.nnz), setup_s=round(setup, 2), newton_wall_s=round(wall, 3), **stats,
         factor_ms_last=round(be.timings()["factor_ms"], 3), solve_ms_last=round(be.timings()["solve_ms"], 3),
         grad_inf=float(np.max(np.abs(g))), posterior_std_s=round(t_std, 3), std_range=[float(sd.min()), float(sd.max())],
         roofline=roofline(be.info(), be.timings()["factor_ms"], be.timings()["solve_ms"]))


def config3():
    cells = 64 if small else 316
    npts = 16 if small else 256
    coords, tri = spde.mesh2d(cells)
    model = spde.MaternSPDE(coords, tri, 1)
    n = model.n
    z = np.random.default_rng(2).standard_normal(n)
    side = int(round(np.sqrt(npts)))
    thetas = [(t, r) for t in np.logspace(-1, 1, side) for r in np.logspace(-1.3, 0, side)]
    Q0 = model.precision(1.0, 0.3)
    from gmrf_b200 import _lib
    lanes = 16
    _lib.set_option("lanes", lanes)
    be = B200Backend(Q0, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0)
    _lib.set_option("lanes", 1)
    basis = model.basis()
    be.set_value_basis(basis)
    zBz = np.array([z @ (sp_mat @ z) for sp_mat in (
        __import__("scipy.sparse", fromlist=["csc_matrix"]).csc_matrix((b, model.rowval, model.colptr), shape=(n, n)) for b in basis)])
    t0 = time.perf_counter()
    dev_ms = 0.0
    out = np.empty(len(thetas))
    for i, (tau, rng_) in enumerate(thetas):
        c = model.coefficients(tau, rng_)
        be.refactorize_combination(c)                 # nzval assembled in HBM, numeric factorization + fused logdet
        dev_ms += be.timings()["factor_ms"]
        quad = float(c @ zBz)                         # z'Q(theta)z through the same basis
        out[i] = 0.5 * be.compute_logdet() - 0.5 * quad - 0.5 * n * np.log(2 * np.pi)
    wall = time.perf_counter() - t0
    # the same sweep, 16 points per launch (lanes)
    coeffs = np.stack([model.coefficients(t, r) for t, r in thetas])
    be.refactorize_combination_lanes(coeffs[:lanes])                 # graph capture outside the timed region
    t1 = time.perf_counter()
    lane_ms = 0.0
    out_l = np.empty(len(thetas))
    for i0 in range(0, len(thetas), lanes):
        c = coeffs[i0:i0 + lanes]
        ld, st = be.refactorize_combination_lanes(c)
        lane_ms += be.timings()["factor_ms"]
        out_l[i0:i0 + lanes] = 0.5 * ld - 0.5 * (c @ zBz) - 0.5 * n * np.log(2 * np.pi)
    wall_l = time.perf_counter() - t1
    # the same points with gradients: + selected inversion + tr(Q^-1 B_j) against the resident basis (d logpdf / d c_j)
    ng = min(16, len(thetas))
    t2 = time.perf_counter()
    sel_ms = 0.0
    grads = np.empty((ng, basis.shape[0]))
    for i, (tau, rng_) in enumerate(thetas[:ng]):
        c = model.coefficients(tau, rng_)
        be.refactorize_combination(c)
        be.selinv_compute(); sel_ms += be.timings()["selinv_ms"]
        grads[i] = 0.5 * (be.selinv_dot_basis() - zBz)
    wall_g = time.perf_counter() - t2
    trace_identity = float(np.max(np.abs(np.array([model.coefficients(*th) @ (2.0 * g + zBz) for th, g in zip(thetas[:ng], grads)]) - n)) / n)
    # spot check one point against the host-assembled matrix
    tau, rng_ = thetas[len(thetas) // 2]
    Q = model.precision(tau, rng_)
    be.refactorize(Q)
    chk = 0.5 * be.compute_logdet() - 0.5 * z @ (Q @ z) - 0.5 * n * np.log(2 * np.pi)
    emit(config=3, n=n, points=len(thetas), sweep_wall_s=round(wall, 3), device_ms_per_eval=round(dev_ms / len(thetas), 3),
         evals_per_s=round(len(thetas) / wall, 1), lanes=lanes, lanes_sweep_wall_s=round(wall_l, 3),
         lanes_device_ms_per_eval=round(lane_ms / len(thetas), 3), lanes_evals_per_s=round(len(thetas) / wall_l, 1),
         lanes_vs_single_max_rel=float(np.max(np.abs(out_l - out) / np.abs(out))), logpdf_spotcheck_rel=float(abs(chk - out[len(thetas) // 2]) / abs(chk)),
         gradient_evals=ng, gradient_wall_ms_per_eval=round(1e3 * wall_g / ng, 3), selinv_ms_per_eval=round(sel_ms / ng, 3),
         trace_identity_rel=trace_identity, status=be.status, roofline={"single": roofline(be.info(), dev_ms / len(thetas)), "lanes": roofline(be.info(), lane_ms / len(thetas))})


def config5():
    cells, nt = (40, 20) if small else (100, 100) if "--big5" in sys.argv else (100, 50)    # --big5: 101^2 x 100 = 1,020,100 latent
    coords, tri = spde.mesh2d(cells)
    model = spde.AdvectionDiffusionSSM(coords, tri, nt=nt)
    rng = np.random.default_rng(3)
    obs = rng.choice(model.ns, min(500, model.ns // 2), replace=False)
    Q = model.posterior(obs, 1.0 / 0.05 ** 2)
    n = Q.shape[0]
    t0 = time.perf_counter()
    # geometric nested dissection with per-axis separator widths (5 hops in space, 1 in time): the fill of METIS (2.57e9 vs
    # 2.71e9 factor entries at 510 k dofs) without its ~25 s of single-threaded ordering; --metis restores the library default
    ordering = None if "--metis" in sys.argv else spde.geometric_nd_perm((cells + 1, cells + 1, nt), leaf=64, width=(5, 5, 1))
    be = B200Backend(Q, ordering=ordering, device=0)
    setup = time.perf_counter() - t0
    info = be.info()
    be.refactorize(Q); f_ms = be.timings()["factor_ms"]
    be.selinv_compute(); s_ms = be.timings()["selinv_ms"]
    var = be.get_selinv_diag()
    m = 1024
    Z = np.asfortranarray(np.random.default_rng(4).standard_normal((m, n)).T)
    X = be.backend_backward_solve(Z); X = be.backend_backward_solve(Z)
    smp_ms = be.timings()["solve_ms"]
    emp = X.var(axis=1)
    emit(config=5, n=n, nnz_q=int(Q.nnz), nnz_l=info["nnz_l"], flops=float(info["flops_chol"]), setup_s=round(setup, 2),
         factor_ms=round(f_ms, 2), selinv_ms=round(s_ms, 2), samples=m, sampling_ms=round(smp_ms, 2),
         ms_per_sample=round(smp_ms / m, 4), var_vs_samples_median_rel=float(np.median(np.abs(emp - var) / var)),
         observed_var_max=float(var[obs].max()),
         roofline={**roofline(info, f_ms, selinv_ms=s_ms),
                   "sampling_tflops": round(2.0 * float(info["nnz_l"]) * m / (smp_ms * 1e-3) / 1e12, 3)})


for c in which:
    {1: config1, 2: config2, 3: config3, 5: config5}[c]()
