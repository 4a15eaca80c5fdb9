#!/bin/bash
# round 2, GPU call 18 (1 GPU): final check of the committed build — whole GPU suite, bench line, 1-D timings,
# compute-sanitizer (memcheck + racecheck) over the small scenarios of both paths
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
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_oned.py -m gpu -q -x -k "steady_nonperiodic or fast_periodic or ramp" > gpurun_out/r02_san_mem_1d.log 2>&1; echo "memcheck 1-D rc=$?"; tail -3 gpurun_out/r02_san_mem_1d.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "matches_oracle_bit_exact and (minimal or tripolar or growing)" > gpurun_out/r02_san_mem_2d.log 2>&1; echo "memcheck 2-D rc=$?"; tail -3 gpurun_out/r02_san_mem_2d.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "matches_oracle_bit_exact and minimal" > gpurun_out/r02_san_race_2d.log 2>&1; echo "racecheck 2-D rc=$?"; tail -3 gpurun_out/r02_san_race_2d.log | cut -c1-200
