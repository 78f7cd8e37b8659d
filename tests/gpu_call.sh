set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=400 run r2o_pytest_ws python -m pytest tests/test_gpu_workspace.py tests/test_gpu_bench_paths.py -m gpu -q -p no:cacheprovider -x
TMO=120 run r2o_plan_2d224 python tests/gpu_plan_profile.py 2d:224 --phase=0 --top=6
TMO=300 run r2o_plan_3d100 python tests/gpu_plan_profile.py 3d:100 --phase=0 --top=8
tail -n 3 gpurun_out/r2o_pytest_ws.log
