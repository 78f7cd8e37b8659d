#!/bin/bash
# One gpurun call's worth of round-opening measurements (each step bounded; everything lands in gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tests/gpu_round_start.sh'
# 1. parity suite  2. smoke  3. where a 2D / the 1 M-dof refactorization spend their time (per launch, GEMM classes by k)
# 4. BASELINE configs 1, 2, 3, 5 with roofline fractions, the few-RHS solve with / without programmatic launches, a 16-lane
#    sweep per launch  5. the headline bench line (parity_check + sharded legs included).
# Steps are independent: a failing step does not stop the rest.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-160))"; }
TMO=600 run pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider
TMO=120 run smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
TMO=120 run plan_profile_2d224 python tests/gpu_plan_profile.py 2d:224 --phase=0,1,2,3 --top=10
TMO=120 run chain_phases python tests/gpu_chain_phases.py 224
TMO=300 run plan_profile_3d100 python tests/gpu_plan_profile.py 3d:100 --phase=0 --top=10
TMO=300 run configs python tests/gpu_configs.py 1 2 3 5
TMO=300 run solve_schedules python tests/gpu_solve_wide.py 3d:100 pdl=0 pdl=1
TMO=200 run plan_profile_lanes python tests/gpu_plan_profile.py 2d:316 --lanes=16 --phase=0 --top=10
TMO=600 run bench python bench.py
tail -n 3 gpurun_out/pytest_gpu.log
