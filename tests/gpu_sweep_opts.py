"""Option sweep on the 2D configs (factor+logdet ms through the CUDA graph): python tests/gpu_sweep_opts.py"""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend
defaults = {"fused_front": 1, "fused_chain": 1, "chain_max_tiles": 160, "asm_gather": 1, "level_alap": 1, "front_smem_kb": 200, "splitk_min_k": 128, "outer_block": 256, "syrk_split": 0}
variants = [{}, {"splitk_min_k": 64}, {"splitk_min_k": 128}, {"splitk_min_k": 256}, {"splitk_min_k": 64, "syrk_split": 1},
            {"splitk_min_k": 128, "syrk_split": 1}, {"splitk_min_k": 256, "syrk_split": 1},
            {"fused_front": 0, "fused_chain": 0, "asm_gather": 0, "level_alap": 0}]
for cells in (224, 316, 500):
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3)
    perm = spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3)
    for v in variants:
        for k, val in {**defaults, **v}.items():
            _lib.set_option(k, val)
        be = B200Backend(Q, ordering=perm, device=0)
        ts = []
        for _ in range(6):
            be.refactorize(Q); ts.append(be.timings()["factor_ms"])
        info = be.info()
        print(f"cells {cells} {str(v):70s} factor {min(ts[1:]):7.3f} ms  launches {info['graph_nodes']:4d} front {info['front_launches']} chain {info['chain_launches']}", flush=True)
        be.close()
for k, val in defaults.items():
    _lib.set_option(k, val)
