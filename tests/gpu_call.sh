set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=900 run r2s_pytest_gpu python -m pytest tests/test_gpu_workspace.py tests/test_cdriver.py -m gpu -q -p no:cacheprovider
tail -n 30 gpurun_out/r2s_pytest_gpu.log | cut -c1-220
