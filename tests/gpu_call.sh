set -u
mkdir -p gpurun_out
T="python tests/gpu_lanes_target.py"
$T > gpurun_out/r2_lanes_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 260 --csv --log-file gpurun_out/r2_launches_lanes.csv $T > gpurun_out/r2_ncu_lanes.log 2>&1
cat gpurun_out/r2_lanes_plain.log | tail -1
python tests/ncu_summary.py gpurun_out/r2_launches_lanes.csv | head -16
