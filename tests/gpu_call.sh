set -u
mkdir -p gpurun_out
T="python tests/gpu_profile_target.py 3d:48 --selinv --solve --order=geo"
$T > gpurun_out/ncu_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none -k regex:assemble_gather_kernel -s 8 -c 3 -f -o gpurun_out/r2_asm $T > gpurun_out/ncu_asm.log 2>&1
ncu --set full --clock-control none -k regex:potrf_inv64_kernel -s 40 -c 2 -f -o gpurun_out/r2_potrf $T > gpurun_out/ncu_potrf.log 2>&1
ncu --set full --clock-control none -k regex:splitk_reduce_kernel -s 2 -c 2 -f -o gpurun_out/r2_splitk $T > gpurun_out/ncu_splitk.log 2>&1
ncu --set full --clock-control none --nvtx --nvtx-include "gmrf_b200:selinv/" -k regex:gemm_dmma_kernel -s 30 -c 8 -f -o gpurun_out/r2_selinv_gemm $T > gpurun_out/ncu_selinv_gemm.log 2>&1
for n in asm potrf splitk selinv_gemm; do python tests/ncu_extract.py gpurun_out/r2_$n.ncu-rep gpurun_out/r2_ncu_$n.txt; rm -f gpurun_out/r2_$n.ncu-rep; done
ls -la gpurun_out/r2_ncu_*.txt
