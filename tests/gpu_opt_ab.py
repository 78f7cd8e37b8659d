"""A/B of one library option on the GPU box (diagnostics, not a pytest file):
    python tests/gpu_opt_ab.py panel_blocked=0,1 2d:224 2d:500 3d:48 [--lanes]
factor+logdet ms through the CUDA graph for every value of the option, log-determinants compared with the first value."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gpu_perf import build_problem  # noqa: E402

key, vals = [a for a in sys.argv[1:] if "=" in a][0].split("=")
vals = [float(v) for v in vals.split(",")]
specs = [a for a in sys.argv[1:] if ":" in a] or ["2d:224"]
for spec in specs:
    Q, dims, width, _ = build_problem(spec)
    ordering = spde.geometric_nd_perm(dims, leaf=64, width=width)
    ref = None
    for v in vals:
        _lib.set_option(key, v)
        be = B200Backend(Q, ordering=ordering, device=0)
        tf = []
        for _ in range(7):
            be.refactorize(Q)
            tf.append(be.timings()["factor_ms"])
        ld = be.compute_logdet()
        x = be.backend_solve(np.ones(Q.shape[0]))
        res = np.linalg.norm(Q @ x - 1.0) / np.sqrt(Q.shape[0])
        if ref is None:
            ref = ld
        print(f"{spec} {key}={v:g}: factor+logdet {min(tf[1:]):8.3f} ms  launches {be.info()['graph_nodes']}  status {be.status}  "
              f"logdet rel dev {abs(ld - ref) / abs(ref):.1e}  solve residual {res:.1e}", flush=True)
        be.close()
if "--lanes" in sys.argv:
    cells, lanes = 316, 16
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3).tocsc()
    for v in vals:
        _lib.set_option(key, v)
        _lib.set_option("lanes", lanes)
        be = B200Backend(Q, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0, factorize=False)
        _lib.set_option("lanes", 1)
        t = []
        for _ in range(4):
            ld, st = be.refactorize_lanes(np.tile(Q.data, (lanes, 1)))
            t.append(be.timings()["factor_ms"])
        print(f"2d:{cells} lanes {lanes} {key}={v:g}: sweep {min(t[1:]):.3f} ms = {min(t[1:]) / lanes:.3f} ms per value set, logdet {ld[0]:.12g}", flush=True)
        be.close()
_lib.set_option(key, vals[-1])
