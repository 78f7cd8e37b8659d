set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --cells 40 > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; echo rc=$?; tail -c 1500 gpurun_out/r2_bench_small.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo rc=$?; tail -c 800 gpurun_out/r2_bench_b.err
