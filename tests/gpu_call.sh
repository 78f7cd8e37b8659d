set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=900 run r2r_pytest_gpu python -m pytest tests -m gpu -q -p no:cacheprovider -x
tail -n 30 gpurun_out/r2r_pytest_gpu.log | cut -c1-200
