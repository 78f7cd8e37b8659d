set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=300 run r2h_pytest_ws python -m pytest tests/test_gpu_workspace.py tests/test_gpu_bench_paths.py -m gpu -q -p no:cacheprovider -x
TMO=120 run r2h_plan_2d224 python tests/gpu_plan_profile.py 2d:224 --phase=0 --top=6
TMO=120 run r2h_plan_2d224_asap python tests/gpu_plan_profile.py 2d:224 --phase=0 --top=6 --set:level_alap=0
TMO=240 run r2h_configs_1_2_3 python tests/gpu_configs.py 1 2 3
TMO=400 run r2h_bench python bench.py --steps 3 --warmup 3 --no-sharded --no-parity
tail -n 3 gpurun_out/r2h_pytest_ws.log
