#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 sparse-Cholesky hot path (BASELINE.json metric).

Workload (N = 1, default): BASELINE.json configs[3] -- 3D Matern SPDE (smoothness 0 => alpha = 2) on a structured
100^3-cell tetrahedral mesh of [-1,1]^3, n = 1,030,301 latent dofs, nnz(Q) = 65.0 M; one STEP = one numeric
refactorization (values -> supernodal LL^T) with the fused log-determinant, same sparsity pattern, new
hyperparameters (tau, range) every step. Selected inversion (marginal variances) is timed separately and reported
in `selinv_ms`.  N > 1: the path does not shard a single factorization (north star: "a single factorization stays
on one GPU"); the ranks evaluate independent hyperparameter points of the same model (one handle per GPU, same
symbolic analysis) and all-gather the log-determinants over NCCL -- weak scaling, no data-path collective.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells 100] [--selinv-reps 1]

`value`      factorizations (+logdet) per second, nzval resident in HBM (gmrf_b200_refactorize_device), whole job.
`e2e`        same through the public API (GMRFWorkspace.update_precision_values + logdet) from HOST buffers:
             host copy into ws.Q.nzval, H2D of nnz(Q) doubles, factorization, D2H of the scalar.
`roofline`   the DMMA GEMM kernel (dominant): algorithmic GEMM flops / summed launch time (CUDA events around every
             launch of one extra refactorization) against the measured cuBLAS FP64 peak of this pool's B200.
`cpu_baseline` multithreaded BLAS-3 supernodal Cholesky on the host cores (oracle/cpu_baseline.py, "port"): bounded
             sample (a smaller mesh of the same recipe), scaled to the full problem by the algorithmic flop count.
--impl reference  times that CPU port alone (all host threads) and prints the same JSON line with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "Cholesky+logdet/s, 1M-dof 3D Matern SPDE Q (selinv ms and FP64 TFLOP/s alongside)"
UNIT = "factorizations/s"
FP64_PEAK_TFLOPS = 36.09      # cuBLAS DGEMM 16384^3 measured on this pool's B200: profiles/r01_fp64_probe.json
FP64_PEAK_SOURCE = ("profiles/r01_fp64_probe.json (cuBLAS DGEMM 16384^3, this pool's B200); MEASURED_PEAKS.json "
                    "carries only HBM and bf16, so the FP64 denominator is this repo's own probe")


def _peak():
    p = os.path.join(ROOT, "profiles", "r01_fp64_probe.json")
    try:
        with open(p) as f:
            return float(json.load(f)["cublas_dgemm_nt_16384_tflops"])
    except Exception:
        return FP64_PEAK_TFLOPS


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def gemm_traffic():
    """DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def theta_for(step: int, rank: int):
    """Hyperparameter point (tau, range) of the sweep evaluated at (step, rank): a 16x16 log grid."""
    i = (step * 7 + rank * 3) % 16
    j = (step * 5 + rank * 11) % 16
    tau = 10.0 ** (-1.0 + 2.0 * i / 15.0)
    rng = 10.0 ** (-1.3 + 1.3 * j / 15.0)
    return tau, max(rng, 0.12)


def build_model(cells: int):
    from gmrf_b200 import spde
    coords, tets = spde.mesh3d(cells)
    model = spde.MaternSPDE(coords, tets, smoothness=0)
    perm = spde.geometric_nd_perm((cells + 1,) * 3, leaf=64, width=2)
    return model, perm


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self._halt = threading.Event()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                out = [o.strip() for o in out]
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for nm, v in zip(names, out[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline_run(sample_cells: int, full_flops: float, reps: int = 1):
    """Time the CPU supernodal port on a smaller mesh of the same recipe; scale to the full problem by flops."""
    import oracle  # noqa: F401  (test/bench infrastructure -- the only place bench.py touches oracle/)
    from oracle.cpu_baseline import CpuSupernodalCholesky
    from gmrf_b200 import _lib
    from gmrf_b200.backend import _Handle
    from gmrf_b200.introspect import Tables
    model, perm = build_model(sample_cells)
    Q = model.precision(*theta_for(0, 0))
    h = _Handle(Q.shape[0], Q.indptr, Q.indices, perm, _lib.ORDER_ND, device=-1)   # host-only symbolic analysis
    T = Tables(h)
    cpu = CpuSupernodalCholesky(T)
    cpu.refactorize(Q.data)                       # warm-up (page faults, thread pools)
    best = min(cpu.refactorize(model.values(*theta_for(r + 1, 0))) for r in range(max(1, reps)))
    flops = float(T.info["flops_chol"])
    gflops = flops / best / 1e9
    value = (gflops * 1e9) / full_flops
    sample = (f"3D Matern alpha=2, {sample_cells}^3 cells (n={Q.shape[0]}, {flops:.3g} flop) factorized in {best:.2f} s "
              f"= {gflops:.0f} GFLOP/s on {cpu.threads} threads; scaled to the {full_flops:.3g}-flop workload")
    h.close()
    return {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port", "sample": sample,
            "gflops": gflops, "sample_seconds": best, "same_config": False,
            "note": "bounded sample scaled by flops (extrapolated); `bench.py --impl reference` measures the full problem",
            "full_problem_measured": _recorded_reference()}


def _recorded_reference():
    """The committed record of the reference arm on this pool's box (whole 100^3 factorizations, measured -- not scaled)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        return {"value": d["value"], "unit": d["unit"], "seconds_per_factorization": d["ms_per_step"] / 1e3,
                "cores": d["cpu_baseline"]["cores"], "source": "profiles/r02_bench_reference_arm.json"}
    except Exception:
        return None


def full_flops_estimate(cells: int) -> float:
    """Algorithmic flop count of the full workload without a GPU (host-only symbolic analysis)."""
    from gmrf_b200 import _lib
    from gmrf_b200.backend import _Handle
    model, perm = build_model(cells)
    h = _Handle(model.n, model.colptr, model.rowval, perm, _lib.ORDER_ND, device=-1)
    fl = float(h.info()["flops_chol"])
    h.close()
    return fl


def _cpu_port(cells):
    """(model, analysis-only handle, tables, CPU port) of the 3D Matern problem at `cells`^3 cells. The symbolic tables come
    from the library's HOST analysis (no GPU is touched: device = -1); every floating-point operation is the port's."""
    import oracle  # noqa: F401
    from oracle.cpu_baseline import CpuSupernodalCholesky
    from gmrf_b200 import _lib
    from gmrf_b200.backend import _Handle
    from gmrf_b200.introspect import Tables
    model, perm = build_model(cells)
    h = _Handle(model.n, model.colptr, model.rowval, perm, _lib.ORDER_ND, device=-1)
    T = Tables(h)
    return model, h, T, CpuSupernodalCholesky(T)


def _independent_crosscheck(cells=16):
    """The CPU port against an implementation that shares NOTHING with this repo's analysis (SciPy's SuperLU, its own
    MMD ordering): relative difference of the log-determinants on a small mesh of the same recipe."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    model, h, T, cpu = _cpu_port(cells)
    nz = model.values(*theta_for(0, 0))
    cpu.refactorize(nz)
    Q = sp.csc_matrix((nz, model.rowval, model.colptr), shape=(model.n, model.n))
    lu = spl.splu(Q, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options={"SymmetricMode": True})
    ld = float(np.sum(np.log(np.abs(lu.U.diagonal()))) + np.sum(np.log(np.abs(lu.L.diagonal()))))
    h.close()
    return {"cells": cells, "n": int(model.n), "logdet_port": cpu.logdet, "logdet_superlu": ld,
            "rel_diff": abs(cpu.logdet - ld) / abs(ld)}


def run_reference(args):
    """CPU arm: the reference's path (CHOLMOD supernodal Cholesky + logdet) cannot run in this image (no Julia, no
    libcholmod), so the arm times oracle/supernodal_cpu.c -- a BLAS-3 supernodal multifrontal port on all host threads.
    `value` is MEASURED on the benchmark's own problem (the full 100^3-cell matrix, same ordering as the GPU arm): as
    many whole factorizations as fit `--cpu-budget` seconds (at least one), never an extrapolation. The flop-scaled
    figure of a smaller mesh is kept as a clearly named secondary field."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    model, h, T, cpu = _cpu_port(args.cells)
    flops = float(T.info["flops_chol"])
    setup_s = time.perf_counter() - t_start
    # untimed first pass (page faults of the 80 GB of panels + update pool, thread pools), then measured steps
    times = []
    warm = cpu.refactorize(model.values(*theta_for(0, 0)))
    k = 0
    while k < max(1, args.steps) and (k == 0 or time.perf_counter() - t_start + (sum(times) / len(times)) < args.cpu_budget):
        times.append(cpu.refactorize(model.values(*theta_for(k + 1, 0))))
        k += 1
    per = sum(times) / len(times)
    value = 1.0 / per
    solve = None
    try:
        rhs = np.random.default_rng(0).standard_normal(model.n)
        cpu.solve(rhs)
        t_full = min(cpu.solve(rhs)[1] for _ in range(2))
        t_half = min(cpu.solve(rhs, half=True)[1] for _ in range(2))
        lbytes = 8.0 * float(T.info["nnz_l_stored"])
        solve = {"solve_ms": round(1e3 * t_full, 1), "half_solve_ms": round(1e3 * t_half, 1),
                 "solve_GBs": round(2.0 * lbytes / t_full / 1e9, 1), "half_solve_GBs": round(lbytes / t_half / 1e9, 1),
                 "measured_on": "the full problem"}
    except Exception as e:
        solve = {"error": str(e)[:200]}
    h.close()
    # secondary, clearly labelled: a smaller mesh scaled by flops (what round 1 reported as the value) + selinv on it
    secondary = None
    try:
        sc = min(args.cells, args.cpu_sample_cells)
        m2, h2, T2, cpu2 = _cpu_port(sc)
        f2 = float(T2.info["flops_chol"])
        cpu2.refactorize(m2.values(*theta_for(0, 0)))
        t2 = cpu2.refactorize(m2.values(*theta_for(1, 0)))
        ts = cpu2.selinv()
        secondary = {"same_config": False, "sample_cells": sc, "sample_n": int(m2.n), "sample_flops": f2, "sample_seconds": round(t2, 3),
                     "sample_gflops": round(f2 / t2 / 1e9, 1), "extrapolated_value": (f2 / t2) / flops,
                     "selinv_sample_seconds": round(ts, 3), "selinv_over_factor": round(ts / t2, 2),
                     "selinv_ms_extrapolated": round(1e3 * ts * flops / f2, 1)}
        h2.close()
    except Exception as e:
        secondary = {"error": str(e)[:200]}
    try:
        cross = _independent_crosscheck()
    except Exception as e:
        cross = {"error": str(e)[:200]}
    sample = (f"{len(times)} whole numeric Cholesky+logdet of the benchmark's own matrix (3D Matern alpha=2, {args.cells}^3 cells, "
              f"n={model.n}, {flops:.3g} flop) after one untimed pass: {per:.1f} s each = {flops / per / 1e9:.0f} GFLOP/s on {cpu.threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": 1, "ms_per_step": 1000.0 * per, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D Matern SPDE alpha=2, {args.cells}^3-cell tetrahedral mesh, n={(args.cells + 1) ** 3}, numeric Cholesky+logdet",
                   "cpu_path": "oracle/supernodal_cpu.c (BLAS-3 supernodal multifrontal, OpenBLAS from SciPy, OpenMP): the reference's CHOLMOD "
                               "path cannot run here (no Julia / libcholmod in the image)",
                   "same_config": True, "full_size_steps_timed": len(times), "steps_requested": args.steps,
                   "cpu_budget_s": args.cpu_budget, "setup_seconds": round(setup_s, 1), "first_pass_seconds": round(warm, 1),
                   "step_seconds": [round(t, 2) for t in times],
                   "symbolic": "host-side analysis of libgmrf_b200 (device = -1, no GPU work); the numeric path is the port's own",
                   "cpu_scope": "per host: rank 0 alone runs, on all host threads, whatever --gpus says"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_solve": solve,
        "secondary_sample": secondary,
        "independent_crosscheck": cross,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_check(be, model, nz_dev, nz_host, info):
    """Numerical checks on the benchmark's own 1 M-dof factorization (VERDICT r01 #1-iii): size-independent identities at
    north_star's tolerances -- normwise backward error of a solve, logdet(2Q) = logdet(Q) + n log 2
    (test/workspace/test_backend_ordering.jl:61-67), and marginal variances from the selected inversion against
    unit-vector solves at sampled indices."""
    import scipy.sparse as sp
    import torch
    n, nnz = info["n"], nz_host[0].size
    rng = np.random.default_rng(11)
    Q = sp.csc_matrix((nz_host[0], model.rowval, model.colptr), shape=(n, n))
    be.refactorize_device(nz_dev[0].data_ptr(), nnz)
    ld = be.compute_logdet()
    b = rng.standard_normal(n)
    x = be.backend_solve(b)
    qnorm = float(abs(Q).sum(axis=1).max())
    berr = float(np.linalg.norm(Q @ x - b) / (np.linalg.norm(b) + qnorm * np.linalg.norm(x)))
    var = be.get_selinv_diag()
    idx = rng.choice(n, 5, replace=False)
    E = np.zeros((n, 5)); E[idx, np.arange(5)] = 1.0
    ref = be.backend_solve(E)[idx, np.arange(5)]
    var_rel = float(np.max(np.abs(var[idx] - ref) / ref))
    z = rng.standard_normal(n)
    smp = be.backend_backward_solve(z)                       # x = P' L^-T z  =>  x' Q x = z' z
    half_rel = float(abs(smp @ (Q @ smp) - z @ z) / (z @ z))
    two = (2.0 * nz_dev[0]).contiguous()
    be.refactorize_device(two.data_ptr(), nnz)
    ld2 = be.compute_logdet()
    log2_rel = float(abs(ld2 - (ld + n * math.log(2.0))) / abs(ld))
    del two
    out = {"solve_backward_error": berr, "logdet_2Q_identity_rel": log2_rel, "selinv_var_vs_unit_solves_max_rel": var_rel,
           "half_solve_quadratic_form_rel": half_rel, "tolerances": {"backward_error": 1e-10, "logdet": 1e-10, "variances": 1e-8, "half_solve": 1e-9},
           "factor_status": be.status}
    out["ok"] = bool(berr <= 1e-10 and log2_rel <= 1e-10 and var_rel <= 1e-8 and half_rel <= 1e-9 and be.status == 0)
    return out


def _bcast_bytes(blob, dist, dev):
    """Broadcast a byte string from rank 0 (the shared symbolic analysis: one per node, workspace_pool.jl:55-58)."""
    import torch
    n = torch.tensor([len(blob) if blob is not None else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=0)
    t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if blob is not None:
        t.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def sweep_config3(world, rank, local, small=False):
    """BASELINE config 3: 256 (tau, range) log-density evaluations on a 100 k-vertex 2D Matern (alpha = 3), STRONG scaling:
    the 256 points are split in contiguous blocks over the ranks (WorkspacePool semantics, workspace_pool.jl:42-67), every
    rank holds one handle built from ONE shared analysis, values are assembled in HBM from the resident basis, 16 points
    advance per launch (lanes), and one all_gather returns the 256 log-densities."""
    import torch
    import torch.distributed as dist
    from gmrf_b200 import _lib, spde
    from gmrf_b200.backend import B200Backend
    from gmrf_b200.sharding import shard_range, all_gather_blocks
    cells, npts, lanes = (64, 64, 8) if small else (316, 256, 16)
    dev = torch.device("cuda", local)
    t0 = time.perf_counter()
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    n = model.n
    Q0 = model.precision(1.0, 0.3)
    blob = None
    _lib.set_option("lanes", lanes)
    try:
        if rank == 0:
            be = B200Backend(Q0, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=local, factorize=False)
            blob = be.export_analysis()
        if world > 1:
            blob = _bcast_bytes(blob, dist, dev)
        if rank != 0:
            be = B200Backend(Q0, analysis=blob, device=local, factorize=False)
    finally:
        _lib.set_option("lanes", 1)
    basis = model.basis()
    be.set_value_basis(basis)
    import scipy.sparse as sp
    z = np.random.default_rng(2).standard_normal(n)
    zBz = np.array([z @ (sp.csc_matrix((b, model.rowval, model.colptr), shape=(n, n)) @ z) for b in basis])
    side = int(round(math.sqrt(npts)))
    thetas = [(t, r) for t in np.logspace(-1, 1, side) for r in np.logspace(-1.3, 0, side)]
    coeffs = np.stack([model.coefficients(t, r) for t, r in thetas])
    lo, hi = shard_range(len(thetas), world, rank)
    setup_s = time.perf_counter() - t0
    be.refactorize_combination_lanes(coeffs[lo:lo + min(lanes, hi - lo)])       # graph capture outside the timed region

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    barrier()
    t1 = time.perf_counter()
    out = np.empty(hi - lo)
    dev_ms = 0.0
    for i0 in range(lo, hi, lanes):
        c = coeffs[i0:min(i0 + lanes, hi)]
        ld, st = be.refactorize_combination_lanes(c)
        dev_ms += be.timings()["factor_ms"]
        out[i0 - lo:i0 - lo + len(c)] = 0.5 * ld - 0.5 * (c @ zBz) - 0.5 * n * math.log(2 * math.pi)
    allv = all_gather_blocks(out.reshape(-1, 1), len(thetas))[:, 0] if world > 1 else out
    barrier()
    wall = time.perf_counter() - t1
    if world > 1:
        t = torch.tensor([wall, dev_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms = t.tolist()
    # spot check of one point against the host-assembled matrix through the ordinary (single-lane) path
    k = len(thetas) // 2
    Qk = model.precision(*thetas[k])
    be.refactorize(Qk)
    chk = 0.5 * be.compute_logdet() - 0.5 * z @ (Qk @ z) - 0.5 * n * math.log(2 * math.pi)
    info = be.info()
    be.close()
    return {"workload": f"256-point (tau, range) sweep, 2D Matern alpha=3, {cells}^2 cells, n={n} (BASELINE configs[2])" if not small else f"small sweep n={n}",
            "scaling": "strong", "points": len(thetas), "lanes_per_launch": lanes, "evaluations_per_s": len(thetas) / wall,
            "wall_ms": 1e3 * wall, "device_ms_max_rank": dev_ms, "setup_seconds": round(setup_s, 1),
            "flops_per_eval": float(info["flops_chol"]), "fp64_tflops": float(info["flops_chol"]) * len(thetas) / wall / 1e12,
            "spotcheck_rel": float(abs(allv[k] - chk) / abs(chk)),
            "collective": "one all_gather of 256 doubles (NCCL)" if world > 1 else "none",
            "shared_analysis_bytes": len(blob) if blob else 0,
            "limiter": "device time of the lane-batched factorizations (latency-bound 2D fronts); host side = 16 coefficient triples per call"}


def sampling_config5(world, rank, local, small=False, nt=None):
    """BASELINE config 5 at 101^2 x 50 = 510,050 latent dofs (the 2 M-latent size does not fit one B200 in this layout):
    space-time advection-diffusion posterior, 1024 posterior samples. Variant A: rank 0 factorizes, the numeric factor
    travels to the peers by NCCL broadcast out of / into the handles' HBM, every rank draws 1024 / N samples (one blocked
    half solve on device-resident white noise). Variant B: every rank factorizes redundantly (no communication)."""
    import torch
    import torch.distributed as dist
    from gmrf_b200 import spde
    from gmrf_b200.backend import B200Backend
    from gmrf_b200.sharding import broadcast_factor, shard_range
    cells, nt_default, m = (24, 10, 256) if small else (100, 50, 1024)
    nt = nt or nt_default
    dev = torch.device("cuda", local)
    t0 = time.perf_counter()
    coords, tri = spde.mesh2d(cells)
    model = spde.AdvectionDiffusionSSM(coords, tri, nt=nt)
    obs = np.random.default_rng(3).choice(model.ns, min(500, model.ns // 2), replace=False)
    Q = model.posterior(obs, 1.0 / 0.05 ** 2)
    n = Q.shape[0]
    perm = spde.geometric_nd_perm((cells + 1, cells + 1, nt), leaf=64, width=(5, 5, 1))
    blob = None
    if rank == 0:
        be = B200Backend(Q, ordering=perm, device=local, factorize=False)
        blob = be.export_analysis()
    if world > 1:
        blob = _bcast_bytes(blob, dist, dev)
    if rank != 0:
        be = B200Backend(Q, analysis=blob, device=local, factorize=False)
    setup_s = time.perf_counter() - t0
    info = be.info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    lo, hi = shard_range(m, world, rank)
    mine = hi - lo
    g = torch.Generator(device=dev); g.manual_seed(4 + rank)
    Z = torch.randn((mine, n), generator=g, device=dev, dtype=torch.float64)        # column-major n x mine
    X = torch.empty_like(Z)

    def draw():
        be.solve_device(Z.data_ptr(), X.data_ptr(), n, mine, half=True)
        return be.timings()["solve_ms"]

    # ---- variant B first: everybody factorizes ----
    be.refactorize(Q)
    draw()                                   # (plans and graphs of this block width are built outside the timed regions)
    barrier()
    t1 = time.perf_counter()
    be.refactorize(Q)
    fB = be.timings()["factor_ms"]
    draw_ms = draw()
    barrier()
    wall_B = time.perf_counter() - t1
    var = be.get_selinv_diag() if rank == 0 else None
    selinv_ms = be.timings()["selinv_ms"] if rank == 0 else 0.0
    # ---- variant A: rank 0 factorizes, broadcast, sharded draws ----
    barrier()
    t2 = time.perf_counter()
    if rank == 0:
        be.refactorize(Q)
    fA = be.timings()["factor_ms"] if rank == 0 else 0.0
    barrier()                               # the peers wait for the factorization here, not inside the timed broadcast
    if world > 1:                           # (connection set-up of the first broadcast stays outside as well)
        w = torch.zeros(1 << 20, device=dev)
        dist.broadcast(w, src=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    broadcast_factor(be, src=0)
    e1.record()
    torch.cuda.synchronize()
    bc_ms = e0.elapsed_time(e1) if world > 1 else 0.0
    draw_ms_A = draw()
    barrier()
    wall_A = time.perf_counter() - t2
    emp_rel = None
    if rank == 0:
        emp = X.var(dim=0, unbiased=True).cpu().numpy() if mine > 1 else None
        if emp is not None:
            emp_rel = float(np.median(np.abs(emp - var) / var))
    if world > 1:
        t = torch.tensor([wall_A, wall_B, bc_ms, draw_ms_A, draw_ms, fB], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_A, wall_B, bc_ms, draw_ms_A, draw_ms, fB = t.tolist()
    bytes_bc = 8.0 * (info["nnz_l_stored"] + 64 * n)          # panels + inverted diagonal blocks (upper bound)
    be.close()
    return {"workload": f"advection-diffusion space-time posterior, {cells + 1}^2 x {nt} = {n} latent dofs, {m} posterior samples (BASELINE configs[4] at reduced size)",
            "scaling": "strong", "samples": m, "samples_per_rank": mine, "setup_seconds": round(setup_s, 1),
            "factor_ms_rank0": fA if rank == 0 else None, "selinv_ms_rank0": selinv_ms,
            "broadcast": {"ms": bc_ms, "GB": bytes_bc / 1e9, "GBs": (bytes_bc / 1e9) / (bc_ms * 1e-3) if bc_ms > 0 else None,
                          "nvlink_peer_copy_GBs_measured_reference": 770.0},
            "draw_ms_max_rank": draw_ms_A,
            "variant_A_factor_once_broadcast": {"wall_ms": 1e3 * wall_A, "samples_per_s": m / wall_A},
            "variant_B_redundant_factorization": {"wall_ms": 1e3 * wall_B, "samples_per_s": m / wall_B, "factor_ms_max_rank": fB, "draw_ms_max_rank": draw_ms},
            "sample_variance_vs_selinv_median_rel": emp_rel,
            "limiter": "A: the factorization on rank 0 plus the broadcast; B: the factorization every rank repeats -- the draws themselves shard linearly"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from gmrf_b200.workspace import GMRFWorkspace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgmrf_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t0 = time.time()
    model, perm = build_model(args.cells)
    # two hyperparameter points are kept resident and alternated; every step re-factorizes different values
    nz_host = [np.ascontiguousarray(model.values(*theta_for(k, rank))) for k in range(2)]
    Q0 = model.precision(*theta_for(0, rank))
    # ONE symbolic analysis per node (workspace_pool.jl:55-58 "resolve the permutation ONCE"): rank 0 analyses on all host
    # threads (torchrun pins OMP_NUM_THREADS=1; every rank analysing at once doubled the set-up in round 1), the blob
    # travels over NCCL and the peers restore it (gmrf_b200_create_from_analysis)
    if world > 1:
        from gmrf_b200.backend import B200Backend
        blob = None
        if rank == 0:
            import contextlib
            try:
                from threadpoolctl import threadpool_limits
                ctx = threadpool_limits(limits=os.cpu_count() or 1, user_api="openmp")
            except Exception:
                ctx = contextlib.nullcontext()
            with ctx:
                ws = GMRFWorkspace(Q0, ordering=perm, device=local)
            blob = ws.backend.export_analysis()
        blob = _bcast_bytes(blob, dist, torch.device("cuda", local))
        if rank != 0:
            ws = GMRFWorkspace(Q0, analysis=blob, device=local)
        del blob
    else:
        ws = GMRFWorkspace(Q0, ordering=perm, device=local)       # symbolic analysis + first factorization
    be = ws.backend
    info = be.info()
    setup_s = time.time() - t0
    nnz = nz_host[0].size
    nz_dev = [torch.from_numpy(a).cuda() for a in nz_host]
    flops = float(info["flops_chol"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: nzval resident in HBM ------------------------------------------------------------------
    logdets = []
    for w in range(args.warmup):
        be.refactorize_device(nz_dev[w % 2].data_ptr(), nnz)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms = 0.0
    t_wall = time.perf_counter()
    for k in range(args.steps):
        be.refactorize_device(nz_dev[k % 2].data_ptr(), nnz)
        dev_ms += be.timings()["factor_ms"]                     # CUDA events on the handle's stream
        logdets.append(be.compute_logdet())
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop()
    status = be.status
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
        # the only exchange of the sharded sweep: gather the per-point log-determinants
        mine = torch.tensor(logdets, device="cuda", dtype=torch.float64)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
    ms_per_step = dev_ms / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- e2e: public API from host buffers -----------------------------------------------------------------
    for w in range(max(1, args.warmup // 2)):
        ws.update_precision_values(nz_host[w % 2]); ws.logdet()
    barrier()
    t_e = time.perf_counter()
    for k in range(args.steps):
        ws.update_precision_values(nz_host[k % 2])
        ws.logdet()
    barrier()
    e2e_s = time.perf_counter() - t_e
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * args.steps / e2e_s
    h2d_ms = be.timings()["h2d_ms"]

    # ---- selected inversion (marginal variances) ---------------------------------------------------------------
    selinv_ms = None
    if args.selinv_reps > 0:
        ts = []
        for r in range(args.selinv_reps + 1):
            be.refactorize_device(nz_dev[r % 2].data_ptr(), nnz)
            be.selinv_compute()
            ts.append(be.timings()["selinv_ms"])
        selinv_ms = min(ts[1:]) if len(ts) > 1 else ts[0]
        if world > 1:
            t = torch.tensor([selinv_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            selinv_ms = float(t.item())

    # ---- triangular solves (mean / Newton step / sampling): 1 right-hand side and one 64-column block ---------------
    solve = None
    if args.solve_reps > 0:
        n = info["n"]
        rng = np.random.default_rng(7)
        be.refactorize_device(nz_dev[0].data_ptr(), nnz)
        b1 = torch.from_numpy(rng.standard_normal(n)).cuda()
        x1 = torch.empty_like(b1)
        t1, tl = [], []
        for r in range(args.solve_reps + 1):
            be.solve_device(b1.data_ptr(), x1.data_ptr(), n, 1)
            t1.append(be.timings()["solve_ms"])
            be.solve_device(b1.data_ptr(), x1.data_ptr(), n, 1, half=True)
            tl.append(be.timings()["solve_ms"])
        B64 = torch.from_numpy(rng.standard_normal((64, n))).cuda()      # 64 columns, column-major n x 64
        X64 = torch.empty_like(B64)
        t64 = []
        for r in range(2):
            be.solve_device(B64.data_ptr(), X64.data_ptr(), n, 64)
            t64.append(be.timings()["solve_ms"])
        bytes_l = 8.0 * info["nnz_l_stored"]
        bytes_alg = 2 * 8.0 * info["nnz_l"] + 4 * 8.0 * n        # exact nnz(L) once per direction + the vectors
        hbm = hbm_peak()
        solve = {"solve_1rhs_ms": min(t1[1:]), "solve_1rhs_GBs": 2 * bytes_l / min(t1[1:]) / 1e6,
                 "solve_1rhs_frac_of_hbm": 2 * bytes_l / min(t1[1:]) / 1e6 / hbm,
                 "solve_1rhs_frac_of_hbm_algorithmic": bytes_alg / min(t1[1:]) / 1e6 / hbm,
                 "lt_solve_1rhs_ms": min(tl[1:]), "lt_solve_1rhs_GBs": bytes_l / min(tl[1:]) / 1e6,
                 "solve_64rhs_ms": min(t64), "solve_64rhs_tflops": 4.0 * info["nnz_l_stored"] * 64 / min(t64) / 1e9,
                 "hbm_peak_GBs": hbm, "bytes": "2 x 8 x nnz(L stored) per forward+backward sweep (the factor panels, streamed once per direction); _algorithmic: 2 x 8 x nnz(L) + 4 x 8 x n (exact nonzeros, SURVEY 8d)"}
        del B64, X64

    # ---- roofline of the dominant kernel (live CUDA events around every launch of one more refactorization) ------
    prof = be.profile_refactorize()
    gemm_ms, gemm_launches, gemm_flops = prof["ms"]["gemm"], prof["launches"]["gemm"], prof["gemm_flops"]
    peak = _peak()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    traffic = gemm_traffic()
    analysis_ms = be.timings()["analysis_ms"]
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = parity_check(be, model, nz_dev, nz_host, info)
        except Exception as e:       # reporting only: never sink the timing line
            parity = {"ok": False, "error": str(e)[:300]}
    # ---- the two workloads that shard, measured at this N inside the same run ---------------------
    sharded = None
    if not args.no_sharded:
        clocks_dev_bytes = info["device_bytes"]
        del nz_dev
        ws.backend.close()
        torch.cuda.empty_cache()
        sharded = {}
        for name, fn in (("config3_theta_sweep", sweep_config3), ("config5_posterior_samples", sampling_config5)):
            try:
                sharded[name] = fn(world, rank, local, small=args.cells < 64)
            except Exception as e:
                sharded[name] = {"error": str(e)[:300]}
            if world > 1:
                dist.barrier()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_run(min(args.cells, args.cpu_sample_cells), flops)
            except Exception as e:  # the baseline is reporting only; never let it sink the GPU number
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"3D Matern SPDE alpha=2 (smoothness 0), {args.cells}^3-cell tetrahedral mesh of [-1,1]^3, "
                            f"n={info['n']}, nnz(Q)={info['nnz_q']}: numeric supernodal Cholesky + fused logdet per step "
                            f"(BASELINE.json configs[3])",
                "ordering": "geometric nested dissection passed as ordering=perm (host side, once per pattern)",
                "nnz_L": info["nnz_l"], "nnz_L_stored": info["nnz_l_stored"], "flops_per_step": flops,
                "supernodes": info["nsuper"], "levels": info["nlevels"], "max_front": info["max_front"],
                "device_GiB": round(info["device_bytes"] / 2 ** 30, 2),
                "l2": "inputs larger than L2 (nzval 0.5 GB, factor 39 GB stream through the 126 MB L2 every step)",
                "parallelism": "1 factorization per GPU" if world == 1 else
                               f"{world} independent hyperparameter points, one per GPU, all_gather of logdets (NCCL)",
                "setup_seconds": round(setup_s, 1), "analysis_ms": round(analysis_ms, 1),
            },
            "fp64_tflops": flops / (ms_per_step * 1e-3) / 1e12,
            "selinv_ms": selinv_ms,
            "selinv_fp64_tflops_equiv": (2.0 * flops / (selinv_ms * 1e-3) / 1e12) if selinv_ms else None,
            "solves": solve,
            "wall_ms_per_step": wall_ms / args.steps,
            "logdet": logdets[-1], "factor_status": status,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nnz * 8), "d2h_bytes_per_step": 12,
                    "h2d_ms": h2d_ms, "api": "GMRFWorkspace.update_precision_values(nzval_host); logdet(ws)"},
            "gpu_launches": int(args.steps * info["graph_nodes"]),
            "roofline": {"bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA.8x8x4 via mma.sync.m8n8k4.f64)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_capture": traffic,
                         "peak_source": FP64_PEAK_SOURCE,
                         "launches_per_step": gemm_launches, "kernel_ms_per_step": gemm_ms,
                         "share_of_step": gemm_ms / sum(prof["ms"].values()) if sum(prof["ms"].values()) > 0 else None,
                         "other_kernels_ms": {k: v for k, v in prof["ms"].items() if k != "gemm"},
                         # the other two phases of the path against the roofline that bounds each (same numbers as the
                         # top-level `selinv_ms` / `solves` keys, kept here so that they travel with the roofline record)
                         "selinv": ({"bound": "tensor", "ms": selinv_ms, "achieved_equiv_2Fchol": 2.0 * flops / (selinv_ms * 1e-3) / 1e12,
                                     "achieved_executed": 2.2 * flops / (selinv_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                     "frac_executed": (2.2 * flops / (selinv_ms * 1e-3) / 1e12) / peak if peak else None,
                                     "note": "Takahashi recursion executes ~2.2 x F_chol flops (W L21 and G as full products)"}
                                    if selinv_ms else None),
                         "solve_1rhs": ({"bound": "hbm", "ms": solve["solve_1rhs_ms"], "achieved": solve["solve_1rhs_GBs"],
                                         "peak": solve["hbm_peak_GBs"], "unit": "GB/s", "frac": solve["solve_1rhs_frac_of_hbm"],
                                         "frac_algorithmic": solve["solve_1rhs_frac_of_hbm_algorithmic"]} if solve else None)},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity_check": parity,
            "sharded": sharded,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=100, help="cells per axis of the 3D mesh (100 -> 1,030,301 dofs)")
    ap.add_argument("--selinv-reps", type=int, default=1)
    ap.add_argument("--solve-reps", type=int, default=2)
    ap.add_argument("--cpu-sample-cells", type=int, default=56)
    ap.add_argument("--cpu-budget", type=float, default=270.0, help="--impl reference: seconds of wall time for setup + measured full-size steps")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check block of the GPU line")
    ap.add_argument("--no-sharded", action="store_true", help="skip the config-3 / config-5 multi-GPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
