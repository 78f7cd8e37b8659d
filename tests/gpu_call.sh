set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=900 run r2t_pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider -x
TMO=120 run r2t_smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
TMO=300 run r2t_configs python tests/gpu_configs.py 1 2 3 5
TMO=600 run r2t_bench python bench.py
B="python bench.py --steps 1 --warmup 3 --no-sharded --no-parity --no-cpu-baseline --selinv-reps 0 --solve-reps 0"
$B > gpurun_out/r2t_bench_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 8450 -c 2810 --csv --log-file gpurun_out/r2t_launches_bench.csv $B > gpurun_out/r2t_ncu_bench.log 2>&1
python tests/ncu_summary.py gpurun_out/r2t_launches_bench.csv > gpurun_out/r2t_launches_bench_summary.txt 2>&1; head -20 gpurun_out/r2t_launches_bench_summary.txt
tail -n 3 gpurun_out/r2t_pytest_gpu.log; tail -2 gpurun_out/r2t_smoke.log
