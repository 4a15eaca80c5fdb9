#!/bin/bash
# round 2, GPU call 10 (1 GPU): exp table in shared memory (smem) against the global-memory table (new) and the
# committed build (head); GPU suite on the smem build; ncu --set full of k_advance (step 4) with source counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t() {
  PICLES_B200_LIB=$PWD/_exp/lib_$1.so python profiles/prof_step.py 4096 12 > gpurun_out/r02_v10_$1.log 2>&1
  python - "$1" <<'PY'
import ast, re, sys
n = sys.argv[1]
ms = []
for line in open(f"gpurun_out/r02_v10_{n}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4:
        ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(f"{n:12s} C2 ms_advance steps 4-11: mean {sum(ms) / len(ms):.4f}  min {min(ms):.4f}" if ms else f"{n}: no data")
PY
}
for v in head new smem new smem; do t $v; done 2>&1 | tee gpurun_out/r02_variants10.txt
for v in new smem; do
  PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/bench_configs.py --steps 5 --warmup 3 --only "C3 growing/decaying winds 2048x2048, on" > gpurun_out/r02_v10_c3_$v.jsonl 2> gpurun_out/r02_v10_c3_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_v10_c3_$v.jsonl').readline()); print('$v C3 ms_advance %.4f prj %.4f'%(d['ms_advance'],d['ms_project_remesh']))" | tee -a gpurun_out/r02_variants10.txt
done
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputests10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests10.log
ncu --set full --clock-control none --import-source on -k regex:'k_advance' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_adv_smem python profiles/prof_step.py 4096 5 > gpurun_out/r02_ncu_adv_smem.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
