#!/bin/bash
# Build the advance kernel with each profiles/ switch of physics.h and time them side by side.
#
#   here (no GPU):   bash profiles/variants.sh build            -> _exp/lib_<name>.so (git-ignored, travels with gpurun)
#   on the B200:     bash profiles/variants.sh time [names...]  -> mean / min ms_advance over steps 4..11 of
#                                                                  prof_step.py 4096 12, one line per variant
# PROF=prof_step_auto.py times the AutoTsit5 instantiation instead.  A variant that wins still has to pass
# `pytest -m gpu` as the default build before it is kept: the host build only proves the formulas, not the
# device fast paths.
set -e
cd "$(dirname "$0")/.."
declare -A V=(
  [base]=""
  [check_all]="-DPH_CHECK_ALL"              # every division of the right-hand side with its validity test (what the unchecked ones buy)
  [vote_per_rhs]="-DPH_VOTE_PER_RHS"        # the steady-wind warp vote at the head of every right-hand side instead of once per particle
)
case "$1" in
build)
  mkdir -p _exp
  for n in "${!V[@]}"; do
    python - "$n" ${V[$n]} <<'PY'
import sys, build_lib
name, flags = sys.argv[1], tuple(sys.argv[2:])
build_lib.build(extra=flags, out=f"_exp/lib_{name}.so")
print("built", name, " ".join(flags))
PY
  done ;;
time)
  shift
  names=("$@"); [ ${#names[@]} -eq 0 ] && names=(base check_all vote_per_rhs base)
  mkdir -p gpurun_out
  for n in "${names[@]}"; do
    PICLES_B200_LIB=$PWD/_exp/lib_$n.so python profiles/${PROF:-prof_step.py} 4096 12 > gpurun_out/variant_$n.log 2>&1
    python - "$n" <<'PY'
import ast, re, sys
n = sys.argv[1]
ms = []
for line in open(f"gpurun_out/variant_{n}.log"):
    m = re.match(r"(\d+) (\{.*\})", line)
    if m and int(m.group(1)) >= 4:
        ms.append(ast.literal_eval(m.group(2))["ms_advance"])
print(f"{n:18s} ms_advance steps 4-11: mean {sum(ms) / len(ms):.4f}  min {min(ms):.4f}")
PY
  done ;;
*) echo "usage: $0 build | time [names...]"; exit 2 ;;
esac
