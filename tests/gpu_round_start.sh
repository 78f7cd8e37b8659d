#!/bin/bash
# One gpurun call's worth of round-opening measurements (each step bounded; everything lands in gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/gpu_round_start.sh'
# 1. parity suite  2. where a 2D refactorization spends its time  3. BASELINE configs 1-3 with roofline fractions
# 4. trace kernels  5. the headline bench line.  Steps are independent: a failing step does not stop the rest.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-160))"; }
TMO=240 run pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider
TMO=120 run plan_profile_2d224 python tests/gpu_plan_profile.py 2d:224 --phase=0 --top=25
TMO=120 run plan_profile_2d500 python tests/gpu_plan_profile.py 2d:500 --phase=0 --top=25
TMO=240 run configs_1_2_3 python tests/gpu_configs.py 1 2 3
TMO=90  run traces_timing python tests/gpu_traces_timing.py 316 5
TMO=420 run bench python bench.py --steps 3 --warmup 3
tail -n 3 gpurun_out/pytest_gpu.log
