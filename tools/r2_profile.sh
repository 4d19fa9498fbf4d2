# round-2 evidence run A (one gpurun call, one ncu use): kernel-level numbers, then the ncu launch list of the bench command
set -x
python tools/kernel_bench.py --cases c2,c1,c3,c4 > gpurun_out/r2_kernel_bench.jsonl 2>&1
python tools/refine_phase_timing.py > gpurun_out/r2_phase_timing.jsonl 2>&1
python tools/eval_profile.py > gpurun_out/r2_eval_sections.jsonl 2>&1
CMD="python bench.py --steps 12 --warmup 3 --cpu-budget 0 --blocks="
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
ls -la gpurun_out/ | tail -12
