"""Phase timing (SM clocks) inside the fused chain-step kernel, last chain launch of a 2D refactorization (diagnostics):
   python tests/gpu_chain_phases.py [cells]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde, _lib
from gmrf_b200.backend import B200Backend
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 224
model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
Q = model.precision(1.0, 0.3)
b = B200Backend(Q, ordering=spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3), device=0)
for _ in range(2):
    st = np.zeros(24, dtype=np.int64)
    b._hd.check(_lib.lib().gmrf_b200_debug_chain_phases(b._hd._h, _lib.ptr(st), 24))
d = np.diff(st[:7])
names = ["load tiles", "panel 0 (192x64)", "store X0", "rank-64 updates", "panel 1 (128x64)", "store X1 / park"]
for nm, v in zip(names, d):
    print(f"{nm:20s} {int(v):8d} cycles  {v / 1.965e3:7.2f} us")
print("total", int(st[6] - st[0]), "cycles")
p = st[8:20]
if p[5] > 0:
    print("inside panel 0 (blocked): sub-panel 0 = 4 rank-4 steps %d cycles, write-back %d, rank-16 DMMA update %d" % (p[6] - p[5], p[7] - p[6], p[8] - p[7]))
    print("   step 1 seen by its look-ahead thread: tile solves + barrier %d, rank-4 update %d, next pivot tile %d, barrier %d  (step %d cycles)"
          % (p[1] - p[0], p[2] - p[1], p[3] - p[2], p[4] - p[3], p[4] - p[0]))
    if p[9] > 0:
        print("   first 32 x 16 warp tile of the rank-16 update: DMMA product %d cycles, C read-modify-write %d" % (p[10] - p[9], p[11] - p[10]))
