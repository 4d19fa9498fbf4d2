# round-2 evidence run (one gpurun call): tests of the changed kernel, kernel-level numbers, bench, then the two ncu captures
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py -q -x -k "refine or fast_path" 2>&1 | tail -3 > gpurun_out/r2_p_tests.log
python tools/kernel_bench.py --cases c2,c1,c4 > gpurun_out/r2_kernel_bench.jsonl 2>&1
python tools/refine_phase_timing.py > gpurun_out/r2_phase_timing.jsonl 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
CMD="python bench.py --steps 12 --warmup 3 --cpu-budget 0 --blocks="
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bf_refine_kernel -s 60 -c 4 -o gpurun_out/r2_refine_bench $CMD > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out/ | tail -20
