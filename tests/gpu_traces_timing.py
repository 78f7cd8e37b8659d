"""GPU timing target (not a test): device-side traces against the selected inverse vs the extract + host-dot route.
Usage: python tests/gpu_traces_timing.py [cells]     (2D Matern alpha = 3 on a cells x cells mesh; default 224 = config 1)
Run plain for wall-clock numbers; run under `ncu -k regex:gather_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` for the kernels' own durations and DRAM traffic."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from gmrf_b200 import spde  # noqa: E402
from gmrf_b200.backend import B200Backend  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 224
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    t0 = time.perf_counter()
    model = spde.MaternSPDE(*spde.mesh2d(cells), 1)
    Q = model.precision(1.0, 0.3)
    basis = model.basis()
    t1 = time.perf_counter()
    be = B200Backend(Q, device=0)
    be.set_value_basis(basis)
    be.refactorize_combination(model.coefficients(1.0, 0.3))
    be.get_selinv_diag()                                     # runs the selected inversion once
    t2 = time.perf_counter()
    out = {"cells": cells, "n": model.n, "nnz_q": int(Q.nnz), "nbasis": int(basis.shape[0]),
           "host_setup_s": round(t1 - t0, 2), "gpu_setup_s": round(t2 - t1, 2)}

    def best(f):
        f()
        ts = []
        for _ in range(reps):
            a = time.perf_counter()
            f()
            ts.append(time.perf_counter() - a)
        return 1e3 * min(ts)

    mats = [sp.csc_matrix((basis[j], Q.indices, Q.indptr), shape=Q.shape) for j in range(basis.shape[0])]
    tr = be.selinv_dot_basis()
    out["basis_traces_ms"] = best(be.selinv_dot_basis)                    # all nbasis traces, resident inputs
    out["selinv_dot_one_matrix_ms"] = best(lambda: be.selinv_dot(mats[1]))  # pattern + values uploaded per call
    out["extract_plus_host_dot_one_matrix_ms"] = best(lambda: float(np.dot(be.selinv_extract_at(mats[1]).data, mats[1].data)))
    host = np.array([float(np.dot(be.selinv_extract_at(m).data, m.data)) for m in mats])
    out["max_rel_diff_vs_host_route"] = float(np.max(np.abs(tr - host) / np.abs(host)))
    out["trace_identity_err"] = float(abs(model.coefficients(1.0, 0.3) @ tr - model.n) / model.n)
    print(json.dumps(out))
    be.close()


if __name__ == "__main__":
    main()
