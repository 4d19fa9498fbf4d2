set -x
python tools/refine_phase_timing.py > gpurun_out/r2_phase1.jsonl 2>&1
for shp in "4,256,2,1" "8,256,2,1" "16,256,2,1" "2,256,2,1" "8,256,1,1" "8,256,2,0"; do
  echo "SHAPE $shp" >> gpurun_out/r2_sweep1.txt
  BF_REFINE_SHAPE=$shp python tools/kernel_bench.py --cases c4 >> gpurun_out/r2_sweep1.txt 2>&1
done
python tools/kernel_bench.py --cases c2,c1,c4 >> gpurun_out/r2_sweep1.txt 2>&1
python -m pytest tests/test_gpu_parity.py -q -x -k refine 2>&1 | tail -3 >> gpurun_out/r2_sweep1.txt
