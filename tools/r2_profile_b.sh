# round-2 evidence run B (one gpurun call, one ncu use): `ncu --set full` of the refinement kernel inside the bench command
set -x
CMD="python bench.py --steps 12 --warmup 3 --cpu-budget 0 --blocks="
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bf_refine_kernel -s 60 -c 4 -o gpurun_out/r2_refine_bench $CMD > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out/ | tail -8
