# quick GPU check of a kernel change: refine / engine parity tests, the bench line without the CPU legs, the C1 / C4 kernel cases
python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -q -x -k "refine or engine" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 --blocks=c1,c5 --cpu-budget 0 > gpurun_out/r2_bench6.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench6.json"))
print("e2e", d["e2e"]["ms_per_step"], "engine", d["e2e_engine"]["ms_per_step"], "value", d["ms_per_step"], "c1", d["c1_step"]["ms_median_of_100"], "roof", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], "cbc", d["e2e_call_by_call"]["ms_per_step"], "c5", d.get("c5", {}).get("keyframes_per_s"))
PY
python tools/kernel_bench.py --cases c2,c1,c4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'case' in d: print(d['case'], d.get('launch'), d['ms'], d.get('iters_mean'))
"
python tools/refine_phase_timing.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['case'], {k: round(v) for k, v in d['cycles_mean'].items() if v})
"
