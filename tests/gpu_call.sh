set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; ( time timeout "$TMO" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 4 gpurun_out/$name.log | head -n 1 | cut -c1-200))"; }
TMO=300 run r2k_pytest_ws python -m pytest tests -m gpu -q -p no:cacheprovider -x
TMO=120 run r2k_chain_phases python tests/gpu_chain_phases.py 224
TMO=120 run r2k_plan_2d224 python tests/gpu_plan_profile.py 2d:224 --phase=0 --top=6
TMO=120 run r2k_plan_2d500 python tests/gpu_plan_profile.py 2d:500 --phase=0 --top=12
TMO=240 run r2k_configs_1_2_3 python tests/gpu_configs.py 1 2 3
tail -n 3 gpurun_out/r2k_pytest_ws.log
