#!/bin/bash
# round 2, GPU call 9 (1 GPU): table-driven exp + unchecked divisions of the advance kernel against the
# committed build (head), on C2 and C3; then the whole GPU suite on the new build (the parity gate)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
t() {  # $1 = variant name
  PICLES_B200_LIB=$PWD/_exp/lib_$1.so python profiles/prof_step.py 4096 12 > gpurun_out/r02_v9_$1.log 2>&1
  python - "$1" <<'PY'
import ast, re, sys
n = sys.argv[1]
ms = []
for line in open(f"gpurun_out/r02_v9_{n}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4:
        ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(f"{n:12s} C2 ms_advance steps 4-11: mean {sum(ms) / len(ms):.4f}  min {min(ms):.4f}" if ms else f"{n}: no data")
PY
}
for v in head checkall new tabconst head new; do t $v; done 2>&1 | tee gpurun_out/r02_variants9.txt
for v in head new tabconst; do
  PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/bench_configs.py --steps 5 --warmup 3 --only "C3 growing/decaying winds 2048x2048, on" > gpurun_out/r02_v9_c3_$v.jsonl 2> gpurun_out/r02_v9_c3_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_v9_c3_$v.jsonl').readline()); print('$v C3 ms_advance %.4f prj %.4f'%(d['ms_advance'],d['ms_project_remesh']))" | tee -a gpurun_out/r02_variants9.txt
done
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputests9.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gputests9.log
