# round-2 evidence run C (one gpurun call, one ncu use): `ncu --set full` of the saturated instantiation on one C4 launch
set -x
CMD="python tools/kernel_bench.py --cases c4f"
$CMD > gpurun_out/r2_c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bf_refine_kernel -s 3 -c 1 -o gpurun_out/r2_refine_c4 $CMD > gpurun_out/r2_ncu_c4.log 2>&1
ls -la gpurun_out/ | grep c4
