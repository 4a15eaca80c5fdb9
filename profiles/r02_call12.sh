#!/bin/bash
# round 2, GPU call 12 (1 GPU): vote once per particle + compile-time propagation + unconditional stage time (final)
# against vote-only and the committed build; DP5 settings; GPU suite on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t() {
  PICLES_B200_LIB=$PWD/_exp/lib_$1.so python profiles/${2:-prof_step.py} 4096 12 > gpurun_out/r02_v12_$1.log 2>&1
  python - "$1" "${2:-prof_step.py}" <<'PY'
import ast, re, sys
n = sys.argv[1]
ms = []
for line in open(f"gpurun_out/r02_v12_{n}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4:
        ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(f"{n:12s} {sys.argv[2]:18s} ms_advance steps 4-11: mean {sum(ms) / len(ms):.4f}  min {min(ms):.4f}" if ms else f"{n}: no data")
PY
}
for v in smem vote final vote final; do t $v; done 2>&1 | tee gpurun_out/r02_variants12.txt
for v in smem vote final; do t $v prof_step_dp5.py; done 2>&1 | tee -a gpurun_out/r02_variants12.txt
for v in smem vote final; do t $v prof_step_auto.py; done 2>&1 | tee -a gpurun_out/r02_variants12.txt
for v in smem final; do
  PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/bench_configs.py --steps 5 --warmup 3 --only "C" > gpurun_out/r02_v12_cfg_$v.jsonl 2> gpurun_out/r02_v12_cfg_$v.err
  python -c "
import json,sys
for l in open('gpurun_out/r02_v12_cfg_$v.jsonl'):
    d=json.loads(l); print('$v', d['config'][:34], 'adv %.4f prj %.4f  %.3e'%(d['ms_advance'],d['ms_project_remesh'],d['particle_steps_per_s']))" | tee -a gpurun_out/r02_variants12.txt
done
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputests12.log 2>&1; echo "pytest rc=$?"; head -4 gpurun_out/r02_gputests12.log | cut -c1-150
