#!/bin/bash
# round 2, GPU call 18 (1 GPU): final check of the committed build — whole GPU suite, bench line, 1-D timings,
# (compute-sanitizer over the small scenarios was planned too: closed on this pool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_gputests18.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests18.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1_final.json"))
print("value %.4e ms %.3f frac %.4f e2e %.4e store %.4e mesh %.4e launches %d cpu %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e_store"]["value"],d["e2e_wind_mesh"]["value"],d["gpu_launches"],d["cpu_baseline"]["value"]))
PY
python profiles/bench_oned.py > gpurun_out/r02_oned.jsonl 2> gpurun_out/r02_oned.err; echo "oned rc=$?"; cat gpurun_out/r02_oned.jsonl | cut -c1-260
# (compute-sanitizer is closed on this pool: the three memcheck / racecheck lines that stood here answered rc=86 and ran nothing)
