"""Target for compute-sanitizer (memcheck / racecheck): every kernel family on small inputs, graphs off.
   compute-sanitizer --tool memcheck python tests/gpu_sanitizer_target.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend

_lib.set_option("use_graph", 0)
rng = np.random.default_rng(0)
cases = [("3d:12 default", spde.MaternSPDE(*spde.mesh3d(12), 0), {}),
         ("2d:48 default (front + chain + gather extend-add)", spde.MaternSPDE(*spde.mesh2d(48), 1), {}),
         ("2d:48 chain everywhere", spde.MaternSPDE(*spde.mesh2d(48), 1), {"fused_front": 0, "chain_max_tiles": 100000}),
         ("3d:12 bulk path, scatter extend-add, split-K", spde.MaternSPDE(*spde.mesh3d(12), 0),
          {"fused_front": 0, "fused_chain": 0, "asm_gather": 0, "splitk_min_k": 32})]
defaults = {"fused_front": 1, "fused_chain": 1, "chain_max_tiles": 160, "asm_gather": 1, "splitk_min_k": 128}
for name, model, opts in cases:
    for k, v in {**defaults, **opts}.items():
        _lib.set_option(k, v)
    Q = model.precision(0.9, 0.5)
    n = Q.shape[0]
    be = B200Backend(Q, device=0)
    be.refactorize(Q)
    x = be.backend_solve(rng.standard_normal(n))
    X = be.backend_solve(rng.standard_normal((n, 70)))
    s = be.backend_backward_solve(rng.standard_normal((n, 3)))
    d = be.get_selinv_diag()
    e = be.selinv_extract_at(Q)
    t = be.selinv_dot(Q)
    info = be.info()
    print(f"{name}: n={n} logdet={be.compute_logdet():.9g} tr={t:.9g} launches={info['graph_nodes']} "
          f"front={info['front_launches']} chain={info['chain_launches']}", flush=True)
    be.close()
print("sanitizer target done")
