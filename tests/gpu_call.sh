set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=900 run r2q_pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider -x
TMO=300 run r2q_configs python tests/gpu_configs.py 1 2
TMO=400 run r2q_bench python bench.py --steps 2 --warmup 3 --no-sharded --no-parity --no-cpu-baseline --selinv-reps 0
tail -n 3 gpurun_out/r2q_pytest_gpu.log
