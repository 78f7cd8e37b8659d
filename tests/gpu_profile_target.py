"""Minimal target for ncu: analysis + 2 refactorizations (+ selinv, + solves) on one problem, graphs off so that
every kernel shows up as its own launch.  python tests/gpu_profile_target.py 3d:48 [--selinv] [--solve] [--solve64] [--order=geo]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend
from gpu_perf import build_problem

spec = sys.argv[1]
order = "geo" if "--order=geo" in sys.argv else "nd"
Q, dims, width, _ = build_problem(spec)
_lib.set_option("use_graph", 0)
ordering = spde.geometric_nd_perm(dims, leaf=64, width=width) if order == "geo" else "nd"
b = B200Backend(Q, ordering=ordering, device=0, factorize=False)
for _ in range(2):
    b.refactorize(Q)
print("factor_ms", b.timings()["factor_ms"], "logdet", b.compute_logdet())
if "--selinv" in sys.argv:
    b.selinv_compute()
    print("selinv_ms", b.timings()["selinv_ms"])
if "--solve" in sys.argv:
    x = b.backend_solve(np.ones(Q.shape[0]))
    print("solve_ms", b.timings()["solve_ms"])
if "--solve64" in sys.argv:
    R = np.asfortranarray(np.random.default_rng(0).standard_normal((Q.shape[0], 64)))
    X = b.backend_solve(R)
    print("solve64_ms", b.timings()["solve_ms"], "residual", np.linalg.norm(Q @ X - R) / np.linalg.norm(R))
