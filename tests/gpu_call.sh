set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=900 run r2p_pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider -x
TMO=120 run r2p_smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
TMO=300 run r2p_configs python tests/gpu_configs.py 1 2 3 5
TMO=600 run r2p_bench python bench.py
tail -n 3 gpurun_out/r2p_pytest_gpu.log; tail -2 gpurun_out/r2p_smoke.log
