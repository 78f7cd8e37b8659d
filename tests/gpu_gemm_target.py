"""ncu target: two launches of the library's own FP64 DMMA GEMM at a large shape (64x64 tile, then 128x64 tile)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import _lib
L = _lib.lib()
m, n, k = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 8192, 2048)))
for flags in (0, 16):
    ms = ctypes.c_double()
    assert L.gmrf_b200_bench_gemm(0, 0, 0, flags, m, n, k, 1, ctypes.byref(ms)) == 0
    print(f"tile={'128x64' if flags else '64x64'} m={m} n={n} k={k}: {ms.value:.3f} ms {2.0*m*n*k/ms.value/1e9:.2f} TFLOP/s")
