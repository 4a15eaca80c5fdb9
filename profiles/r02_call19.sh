#!/bin/bash
# round 2, GPU call 19 (1 GPU): AutoTsit5 instantiation with compile-time propagation — its tests and its time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_autotsit5.py tests/test_gpu_golden.py tests/test_gpu_configs.py -m gpu -q -x ) > gpurun_out/r02_gputests19.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests19.log | cut -c1-200
for v in cur; do
  if [ $v = final ]; then export PICLES_B200_LIB=$PWD/_exp/lib_final.so; else unset PICLES_B200_LIB; fi
  python profiles/prof_step_auto.py 4096 12 > gpurun_out/r02_v19_$v.log 2>&1
  python - $v <<'PY'
import ast, re, sys
ms=[]
for line in open(f"gpurun_out/r02_v19_{sys.argv[1]}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4: ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(sys.argv[1], "AutoTsit5 ms_advance steps 4-11: mean %.4f min %.4f" % (sum(ms)/len(ms), min(ms)))
PY
done
