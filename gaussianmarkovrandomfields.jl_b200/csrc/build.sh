#!/bin/bash
# Builds libgmrf_b200.so in-tree (sm_100a only). Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
METIS="$CUDA_HOME/targets/x86_64-linux/lib/libmetis_static.a"
g++ -std=c++17 -O2 -fPIC -fopenmp -c "$HERE/symbolic.cpp" -o "$OUT/symbolic.o"
"$CUDA_HOME/bin/nvcc" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-fopenmp -Xptxas -v "$@" -c "$HERE/gmrf_b200.cu" -o "$OUT/gmrf_b200.o" 2> "$OUT/ptxas.log" || { cat "$OUT/ptxas.log"; exit 1; }
"$CUDA_HOME/bin/nvcc" -shared -o "$OUT/libgmrf_b200.so" "$OUT/gmrf_b200.o" "$OUT/symbolic.o" "$METIS" -Xcompiler -fopenmp -lgomp
echo "built $OUT/libgmrf_b200.so"
