"""Programmatic dependent launch on the factorization / selected-inversion plans (option pdl_factor), off against on:
    python tests/gpu_sweep_pdl.py [2d:224 2d:500 3d:48 ...]
factor+logdet and selinv ms through the CUDA graphs, results compared bit for bit (the kernels and their order do not change)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gpu_perf import build_problem  # noqa: E402

specs = [a for a in sys.argv[1:] if ":" in a] or ["2d:224", "2d:500", "3d:48"]
do_selinv = "--no-selinv" not in sys.argv
for spec in specs:
    Q, dims, width, _ = build_problem(spec)
    ordering = spde.geometric_nd_perm(dims, leaf=64, width=width)
    ref = None
    for pdl in (0, 1):
        _lib.set_option("pdl_factor", pdl)
        be = B200Backend(Q, ordering=ordering, device=0)
        tf, ts = [], []
        for _ in range(6):
            be.refactorize(Q)
            tf.append(be.timings()["factor_ms"])
        ld = be.compute_logdet()
        d = None
        if do_selinv:
            for _ in range(3):
                be.refactorize(Q)
                be.selinv_compute()
                ts.append(be.timings()["selinv_ms"])
            d = be.get_selinv_diag().copy()
        same = "" if ref is None else f"  bit-identical to pdl_factor=0: logdet {ld == ref[0]}" + (f", selinv diag {bool(np.array_equal(d, ref[1]))}" if do_selinv else "")
        if ref is None:
            ref = (ld, d)
        print(f"{spec} pdl_factor={pdl}: factor+logdet {min(tf[1:]):8.3f} ms" + (f"  selinv {min(ts[1:]):9.3f} ms" if ts else "") +
              f"  launches {be.info()['graph_nodes']}{same}", flush=True)
        be.close()
_lib.set_option("pdl_factor", 0)
# lane sweep of config 3 (16 value sets per launch)
if "--lanes" in sys.argv:
    cells, lanes = 316, 16
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3).tocsc()
    for pdl in (0, 1):
        _lib.set_option("pdl_factor", pdl)
        _lib.set_option("lanes", lanes)
        be = B200Backend(Q, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0, factorize=False)
        _lib.set_option("lanes", 1)
        t = []
        for _ in range(4):
            ld, st = be.refactorize_lanes(np.tile(Q.data, (lanes, 1)))
            t.append(be.timings()["factor_ms"])
        print(f"2d:{cells} lanes {lanes} pdl_factor={pdl}: sweep {min(t[1:]):.3f} ms = {min(t[1:]) / lanes:.3f} ms per value set, logdet {ld[0]:.12g}", flush=True)
        be.close()
    _lib.set_option("pdl_factor", 0)
