#!/bin/bash
# round 2, GPU call 6 (1 GPU): whole GPU suite on the current build, bench line, graded upload blocks, base timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_gputests6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gputests6.log
python profiles/prof_step.py 4096 12 > gpurun_out/r02_prof_step.log 2>&1; python - <<'PY'
import ast,re
ms=[]
for line in open("gpurun_out/r02_prof_step.log"):
    m=re.match(r"(\d+) (\{.*\})",line)
    if m and int(m.group(1))>=4: ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print("ms_advance steps 4-11: mean %.4f min %.4f"%(sum(ms)/len(ms),min(ms)))
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench rc=$?"
PICLES_PIPE_FIRST_ROWS=64 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_graded.json 2> gpurun_out/r02_bench_n1_graded.err; echo "bench graded rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n1_b","r02_bench_n1_graded"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value %.4e ms %.3f frac %.4f e2e %.4e (%.3f ms) store %.4e (%.3f ms) mesh %.4e launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e_store"]["value"],d["e2e_store"]["ms_per_step"],d["e2e_wind_mesh"]["value"],d["gpu_launches"]))
    print("  hbm", [(r["kernel"], round(r["frac"],3), round(r["ms_per_launch"],4)) for r in d["roofline_hbm"]], d.get("cpu_baseline",{}).get("value"))
PY
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/r02_bench_ref.json
