#!/bin/bash
# round 2, GPU call 11 (1 GPU): per-stage specialised stage tails (tails), steady-wind vote once per particle (vote),
# both, against the committed build (smem); C2, C3 and the DP5 (bench06) settings; GPU suite on the `both` build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t() {
  PICLES_B200_LIB=$PWD/_exp/lib_$1.so python profiles/${2:-prof_step.py} 4096 12 > gpurun_out/r02_v11_$1.log 2>&1
  python - "$1" "${2:-prof_step.py}" <<'PY'
import ast, re, sys
n = sys.argv[1]
ms = []
for line in open(f"gpurun_out/r02_v11_{n}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4:
        ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(f"{n:12s} {sys.argv[2]:18s} ms_advance steps 4-11: mean {sum(ms) / len(ms):.4f}  min {min(ms):.4f}" if ms else f"{n}: no data")
PY
}
for v in smem tails vote both smem both; do t $v; done 2>&1 | tee gpurun_out/r02_variants11.txt
for v in smem both; do t $v prof_step_dp5.py; done 2>&1 | tee -a gpurun_out/r02_variants11.txt
for v in smem both; do
  PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/bench_configs.py --steps 5 --warmup 3 --only "C3 growing/decaying winds 2048x2048, on" > gpurun_out/r02_v11_c3_$v.jsonl 2> gpurun_out/r02_v11_c3_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_v11_c3_$v.jsonl').readline()); print('$v C3 ms_advance %.4f prj %.4f'%(d['ms_advance'],d['ms_project_remesh']))" | tee -a gpurun_out/r02_variants11.txt
done
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputests11.log 2>&1; echo "pytest rc=$?"; head -5 gpurun_out/r02_gputests11.log | cut -c1-150
