"""Option sweep for the lane-batched sweep of config 3 (16 value sets per launch): python tests/gpu_sweep_lanes.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend
defaults = {"fused_front": 1, "fused_chain": 1, "chain_max_tiles": 160, "asm_gather": 1, "level_alap": 1}
variants = [{}, {"fused_chain": 0}, {"fused_front": 0}, {"fused_front": 0, "fused_chain": 0}, {"asm_gather": 0},
            {"fused_chain": 0, "asm_gather": 0}, {"fused_front": 0, "fused_chain": 0, "asm_gather": 0, "level_alap": 0}]
cells = 316
model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
Q = model.precision(1.0, 0.3)
perm = spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3)
thetas = [(t, r) for t in np.logspace(-1, 1, 8) for r in np.logspace(-1.3, 0, 8)]
coeffs = np.stack([model.coefficients(t, r) for t, r in thetas])
for lanes in (16, 32):
    for v in variants:
        for k, val in {**defaults, **v}.items():
            _lib.set_option(k, val)
        _lib.set_option("lanes", lanes)
        be = B200Backend(Q, ordering=perm, device=0, factorize=False)
        _lib.set_option("lanes", 1)
        be.set_value_basis(model.basis())
        be.refactorize_combination_lanes(coeffs[:lanes])
        ms = 0.0
        for i0 in range(0, len(thetas), lanes):
            be.refactorize_combination_lanes(coeffs[i0:i0 + lanes]); ms += be.timings()["factor_ms"]
        print(f"lanes {lanes} {str(v):75s} {ms / len(thetas):7.3f} ms/eval  {1e3 * len(thetas) / ms:7.1f} evals/s (device)", flush=True)
        be.close()
for k, val in defaults.items():
    _lib.set_option(k, val)
