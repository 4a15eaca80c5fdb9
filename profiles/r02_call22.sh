#!/bin/bash
# round 2, GPU call 22 (1 GPU): what leaving four SMs empty costs the advance kernel by itself (the remedy reverted in
# DESIGN.md §5 combined it with a high-priority communication stream and lost 20 % on two GPUs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in cur reserve4 cur; do
  PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/prof_step.py 4096 10 > gpurun_out/r02_v22_$v.log 2>&1
  python - $v <<'PY'
import ast, re, sys
ms=[]
for line in open(f"gpurun_out/r02_v22_{sys.argv[1]}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4: ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print("%-10s C2 ms_advance steps 4-9: mean %.4f min %.4f" % (sys.argv[1], sum(ms)/len(ms), min(ms)))
PY
done 2>&1 | tee gpurun_out/r02_variants22.txt
