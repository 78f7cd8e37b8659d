"""Export the BASELINE.json config matrices (and the permutation the B200 arm uses) as raw little-endian binaries that
`baseline/cholmod_ref.jl` reads, so the reference's own CHOLMOD path can be timed on identical inputs wherever Julia
exists. Not used by tests or bench.py.

    python baseline/export_configs.py OUTDIR [1] [3] [4] [--small]

File layout (`<name>.bin`): int64 n, int64 nnz, int64 has_perm, then colptr[n+1] (1-based int64), rowval[nnz] (1-based
int64), nzval[nnz] (float64), perm[n] (1-based int64, only if has_perm)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import spde  # noqa: E402


def write(path, Q, perm=None):
    Q = Q.tocsc()
    Q.sort_indices()
    n = Q.shape[0]
    with open(path, "wb") as f:
        np.array([n, Q.nnz, 0 if perm is None else 1], dtype="<i8").tofile(f)
        (Q.indptr.astype("<i8") + 1).tofile(f)
        (Q.indices.astype("<i8") + 1).tofile(f)
        Q.data.astype("<f8").tofile(f)
        if perm is not None:
            (np.asarray(perm, dtype="<i8") + 1).tofile(f)
    print(f"{path}: n={n} nnz={Q.nnz}", flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out = args[0] if args else "baseline/_matrices"
    which = [int(a) for a in args[1:]] or [1, 3, 4]
    small = "--small" in sys.argv
    os.makedirs(out, exist_ok=True)
    for c in which:
        if c in (1, 2, 3):          # 2D Matern alpha = 3: 224 / 500 / 316 cells
            cells = {1: 224, 2: 500, 3: 316}[c] if not small else 48
            Q = spde.MaternSPDE(*spde.mesh2d(cells), 1).precision(1.0, 0.3)
            perm = spde.geometric_nd_perm((cells + 1, cells + 1), leaf=64, width=3)
        elif c == 4:                # 3D Matern alpha = 2, 100^3 cells (the bench workload)
            cells = 100 if not small else 16
            Q = spde.MaternSPDE(*spde.mesh3d(cells), 0).precision(1.0, 0.5)
            perm = spde.geometric_nd_perm((cells + 1,) * 3, leaf=64, width=2)
        else:
            raise SystemExit(f"config {c}: export not implemented")
        write(os.path.join(out, f"config{c}{'_small' if small else ''}.bin"), Q, perm)


if __name__ == "__main__":
    main()
