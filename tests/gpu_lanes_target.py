"""ncu target: one 16-lane refactorization sweep of config 3 (graphs off so that every kernel is its own launch)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend
cells, lanes = 316, 16
model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
Q = model.precision(1.0, 0.3)
_lib.set_option("use_graph", 0)
_lib.set_option("lanes", lanes)
be = B200Backend(Q, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0, factorize=False)
_lib.set_option("lanes", 1)
be.set_value_basis(model.basis())
thetas = [(t, r) for t in np.logspace(-1, 1, 4) for r in np.logspace(-1.3, 0, 4)]
coeffs = np.stack([model.coefficients(t, r) for t, r in thetas])
for _ in range(2):
    ld, st = be.refactorize_combination_lanes(coeffs)
print("factor_ms", be.timings()["factor_ms"], "per eval", be.timings()["factor_ms"] / lanes)
