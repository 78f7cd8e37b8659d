"""Wide right-hand-side sweeps (GEMM sweeps, 64 / 256 columns per pass) with and without programmatic dependent launch:
    python tests/gpu_multi_pdl.py 3d:48 [3d:100]
pdl_multi = 1: the GEMMs of a sweep are launched programmatically and load their first factor tiles (operand A, constant
during a sweep) before waiting for the predecessor."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gpu_perf import build_problem  # noqa: E402

for spec in [a for a in sys.argv[1:] if ":" in a] or ["3d:48"]:
    Q, dims, width, _ = build_problem(spec)
    n = Q.shape[0]
    ordering = spde.geometric_nd_perm(dims, leaf=64, width=width)
    rng = np.random.default_rng(0)
    ref = {}
    for pdl in (0, 1):
        _lib.set_option("pdl_multi", pdl)
        b = B200Backend(Q, ordering=ordering, device=0)
        info = b.info()
        line = f"{spec} pdl_multi={pdl}:"
        for m in (64, 256):
            if n * m * 8 > 3e9:
                continue
            R = np.asfortranarray(np.random.default_rng(m).standard_normal((n, m)))
            ts, tl = [], []
            for _ in range(3):
                X = b.backend_solve(R); ts.append(b.timings()["solve_ms"])
                Y = b.backend_backward_solve(R); tl.append(b.timings()["solve_ms"])
            fl = 4.0 * info["nnz_l_stored"] * m
            same = ""
            if pdl == 0:
                ref[m] = (X, Y)
            else:
                same = f" bit-identical {bool(np.array_equal(X, ref[m][0]) and np.array_equal(Y, ref[m][1]))}"
            line += f"  {m} rhs: solve {min(ts[1:]):.3f} ms ({fl / min(ts[1:]) / 1e9:.2f} TFLOP/s), Lt-solve {min(tl[1:]):.3f} ms, residual {np.linalg.norm(Q @ X - R) / np.linalg.norm(R):.1e}{same};"
        print(line, flush=True)
        b.close()
_lib.set_option("pdl_multi", 0)
