"""Few-RHS solve time for the step schedules (diagnostics, not a pytest file):
    python tests/gpu_solve_wide.py 3d:100 wide_steps=0,pdl=0 wide_steps=0,pdl=1 wide_steps=2,pdl=0
wide_steps 0 = 64-column steps, 1 = 256-column steps + diagonal-block launches, 2 = 256-column steps with a look-ahead head;
pdl 1 = block steps launched with programmatic stream serialization."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402
from gpu_perf import build_problem  # noqa: E402

spec = sys.argv[1] if len(sys.argv) > 1 else "3d:48"
variants = [dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in a.split(",")) for a in sys.argv[2:]] or [{"pdl": 0}, {"pdl": 1}]
defaults = {"wide_steps": 0, "pdl": 1}
Q, dims, width, _ = build_problem(spec)
n = Q.shape[0]
ordering = spde.geometric_nd_perm(dims, leaf=64, width=width)
rng = np.random.default_rng(0)
rhs = rng.standard_normal(n)
R8 = rng.standard_normal((n, 8))
ref = None
for mode in variants:
    for k, v in {**defaults, **mode}.items():
        _lib.set_option(k, v)
    b = B200Backend(Q, ordering=ordering, device=0)
    info = b.info()
    ts, tl, t8 = [], [], []
    for _ in range(5):
        x = b.backend_solve(rhs); ts.append(b.timings()["solve_ms"])
        s = b.backend_backward_solve(rhs); tl.append(b.timings()["solve_ms"])
        X = b.backend_solve(R8); t8.append(b.timings()["solve_ms"])
    res = np.linalg.norm(Q @ x - rhs) / np.linalg.norm(rhs)
    res8 = np.linalg.norm(Q @ X - R8) / np.linalg.norm(R8)
    bytes_solve = 2 * 8 * info["nnz_l_stored"]
    nl = [len(b.profile_plan(p)) for p in (2, 3)]
    dev = "" if ref is None else f"  max rel dev from the first: {np.max(np.abs(x - ref)) / np.max(np.abs(ref)):.1e}"
    if ref is None:
        ref = x
    print(f"{spec} {mode}: solve 1 rhs {min(ts[1:]):.3f} ms ({bytes_solve / min(ts[1:]) / 1e6:.0f} GB/s), Lt-solve {min(tl[1:]):.3f} ms, "
          f"8 rhs {min(t8[1:]):.3f} ms; launches fwd/bwd {nl[0]}/{nl[1]}; residual {res:.1e} / {res8:.1e}{dev}", flush=True)
    b.close()
for k, v in defaults.items():
    _lib.set_option(k, v)
