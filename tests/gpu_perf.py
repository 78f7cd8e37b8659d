"""Performance exploration on the GPU box (not a pytest file):
    python tests/gpu_perf.py 2d:224 3d:32 3d:48 [--selinv] [--order geo|nd]
Times analysis, refactorize (nzval resident in HBM), solves and selected inversion, and checks size-independent
properties (residual, logdet scaling, selinv diagonal against unit-vector solves)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402


def build_problem(spec):
    kind, size = spec.split(":")
    size = int(size)
    t = time.time()
    if kind == "2d":
        c, e = spde.mesh2d(size)
        m = spde.MaternSPDE(c, e, 1)
        dims, width = (size + 1, size + 1), 3
    else:
        c, e = spde.mesh3d(size)
        m = spde.MaternSPDE(c, e, 0)
        dims, width = (size + 1,) * 3, 2
    Q = m.precision(1.0, 0.3)
    return Q, dims, width, time.time() - t


def run(spec, order, do_selinv, reps=3):
    import torch
    Q, dims, width, tb = build_problem(spec)
    n = Q.shape[0]
    t = time.time()
    ordering = spde.geometric_nd_perm(dims, leaf=64, width=width) if order == "geo" else "nd"
    b = B200Backend(Q, ordering=ordering, device=0, factorize=False)
    ta = time.time() - t
    info = b.info()
    print(f"== {spec} order={order}: n={n} nnzQ={Q.nnz} build {tb:.1f}s analysis+upload {ta:.1f}s "
          f"nnzL={info['nnz_l']:.4g} stored={info['nnz_l_stored']:.4g} flops={info['flops_chol']:.4g} "
          f"nsuper={info['nsuper']} levels={info['nlevels']} maxfront={info['max_front']} maxns={info['max_ns']} "
          f"pool={info['update_pool']:.3g} dev={info['device_bytes'] / 2**30:.2f}GiB launches={info['graph_nodes']}", flush=True)
    nz = torch.from_numpy(Q.data).cuda()
    times = []
    for r in range(reps + 1):
        b.refactorize_device(nz.data_ptr(), nz.numel())
        times.append(b.timings()["factor_ms"])
    fms = min(times[1:])
    ld = b.compute_logdet()
    print(f"   factor {fms:.2f} ms (first {times[0]:.2f})  -> {info['flops_chol'] / fms / 1e9:.2f} TFLOP/s (algorithmic), "
          f"{1000 / fms:.1f} fact/s, status={b.status}, logdet={ld:.12g}", flush=True)
    # host-buffer path (e2e)
    t = time.time(); b.refactorize(Q); te = time.time() - t
    print(f"   refactorize from host buffer: {te * 1e3:.2f} ms wall (h2d {b.timings()['h2d_ms']:.2f} ms)")
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal(n)
    x = b.backend_solve(rhs)
    t1 = b.timings()["solve_ms"]
    x = b.backend_solve(rhs)
    t1 = min(t1, b.timings()["solve_ms"])
    res = np.linalg.norm(Q @ x - rhs) / np.linalg.norm(rhs)
    bytes_solve = 2 * 8 * info["nnz_l_stored"]
    print(f"   solve 1 rhs {t1:.2f} ms  ({bytes_solve / t1 / 1e6:.0f} GB/s of L streamed)  rel residual {res:.2e}")
    R = rng.standard_normal((n, 8))
    X = b.backend_solve(R)
    print(f"   solve 8 rhs {b.timings()['solve_ms']:.2f} ms  residual {np.linalg.norm(Q @ X - R) / np.linalg.norm(R):.2e}")
    for m in (64, 256):
        if n * m * 8 > 3e9:
            continue
        R = rng.standard_normal((n, m))
        X = b.backend_solve(R)
        X = b.backend_solve(R)
        tm = b.timings()["solve_ms"]
        fl = 4.0 * info["nnz_l_stored"] * m
        print(f"   solve {m} rhs (GEMM sweeps) {tm:.2f} ms  = {tm / m:.3f} ms/rhs, {fl / tm / 1e9:.2f} TFLOP/s  residual {np.linalg.norm(Q @ X - R) / np.linalg.norm(R):.2e}")
    z = rng.standard_normal(n)
    s = b.backend_backward_solve(z)
    s = b.backend_backward_solve(z)
    print(f"   Lt-solve 1 rhs {b.timings()['solve_ms']:.2f} ms  ({8 * info['nnz_l_stored'] / b.timings()['solve_ms'] / 1e6:.0f} GB/s of L streamed)")
    # logdet(2Q) = logdet(Q) + n log 2
    b.refactorize(Q * 2.0)
    ld2 = b.compute_logdet()
    print(f"   logdet(2Q)-logdet(Q)-n log2 = {ld2 - ld - n * np.log(2.0):.3e} (rel {abs(ld2 - ld - n * np.log(2.0)) / abs(ld2):.1e})")
    b.refactorize(Q)
    if do_selinv:
        t = time.time()
        d = b.get_selinv_diag()
        tw = time.time() - t
        ts = b.timings()["selinv_ms"]
        b.refactorize_device(nz.data_ptr(), nz.numel())
        b.selinv_compute()
        ts2 = b.timings()["selinv_ms"]
        idx = rng.choice(n, 6, replace=False)
        E = np.zeros((n, idx.size)); E[idx, np.arange(idx.size)] = 1.0
        S = b.backend_solve(E)
        ref = S[idx, np.arange(idx.size)]
        print(f"   selinv {ts2:.2f} ms (first {ts:.2f}, wall incl. alloc {tw * 1e3:.0f} ms) -> {2 * info['flops_chol'] / ts2 / 1e9:.2f} TFLOP/s-equiv; "
              f"diag vs unit solves max rel err {np.max(np.abs(d[idx] - ref) / ref):.2e}", flush=True)
    b.close()
    del nz
    torch.cuda.empty_cache()


def gemm_bench():
    import ctypes
    L = _lib.lib()
    shapes = [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 256, 8192), (16384, 16384, 64), (16384, 192, 64),
              (8192, 8192, 512), (8192, 64, 4096), (2048, 2048, 2048), (20000, 20000, 10000)]
    for (ta, tb, nm) in ((0, 0, "A m-major, B n-major (Cholesky updates)"), (0, 1, "A m-major, B k-major"), (1, 1, "A,B k-major")):
        for large in (16, 0):
            for (m, n, k) in shapes:
                if (m * n * k > 5e11 and (large == 0 or ta)):
                    continue
                ms = ctypes.c_double()
                rc = L.gmrf_b200_bench_gemm(0, ta, tb, large, m, n, k, 3, ctypes.byref(ms))
                assert rc == 0
                print(f"   gemm [{nm}] tile={'128' if large else '64'} m={m} n={n} k={k}: {ms.value:.3f} ms  {2.0 * m * n * k / ms.value / 1e9:.2f} TFLOP/s", flush=True)


def gemm_variants():
    import ctypes
    L = _lib.lib()
    names = {0: "prod 64x64 2x2 KT16 ST3", 1: "128x64 2x2(w64x32) KT16 ST3", 2: "128x128 4x4(w32x32) KT16 ST3", 3: "64x64 2x2 KT32 ST3",
             4: "128x64 4x2(w32x32) KT16 ST3", 5: "64x128 2x4(w32x32) KT16 ST3", 6: "128x128 2x4 KT16 ST4", 7: "64x64 2x2 KT16 ST4",
             8: "128x64 4x2 KT32 ST3", 99: "prod 128x128 2x4 KT16 ST3"}
    for (m, n, k) in [(8192, 8192, 8192), (16384, 256, 8192), (8192, 8192, 256), (16384, 192, 64), (12000, 12000, 64)]:
        for v in (0, 99, 1, 2, 3, 4, 5, 6, 7, 8):
            flags = 16 if v == 99 else (v << 8)
            ms = ctypes.c_double()
            rc = L.gmrf_b200_bench_gemm(0, 0, 0, flags, m, n, k, 3, ctypes.byref(ms))
            assert rc == 0, L.gmrf_b200_last_error(None)
            print(f"   variant {names[v]:32s} m={m} n={n} k={k}: {ms.value:8.3f} ms  {2.0 * m * n * k / ms.value / 1e9:6.2f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if "--gemm-variants" in sys.argv:
        gemm_variants()
    if "--gemm" in sys.argv:
        gemm_bench()
    for a in sys.argv[1:]:
        if a.startswith("--opt="):                      # library tunables, e.g. --opt=bwd_row_chunk=1024
            k, v = a[6:].split("=")
            _lib.set_option(k, float(v))
            print(f"option {k} = {v}")
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    order = "nd"
    for a in sys.argv[1:]:
        if a.startswith("--order"):
            order = a.split("=")[1]
    for spec in args:
        run(spec, order, "--selinv" in sys.argv)
