#!/bin/bash
# The Blackwell proof as a tracked artefact: what the in-tree library is compiled for and which instructions
# carry the claims of DESIGN.md (TMA tile loads + mbarrier in the gather, FP64 FMAs in the advance kernel,
# no local-memory spills in the hot kernels).  No GPU needed:
#     bash profiles/sass_proof.sh > profiles/r02_sass_proof.txt
set -e
cd "$(dirname "$0")/.."
LIB=picles_b200/libpicles_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  commit $(git rev-parse --short HEAD)$(git diff --quiet -- picles_b200/csrc include || echo ' + uncommitted changes')  $LIB"
echo "# cuobjdump -lelf: embedded cubins"
cuobjdump -lelf $LIB
TMP=$(mktemp)
cuobjdump -sass $LIB > $TMP
echo
echo "# per kernel: SASS instructions, FP64 (DFMA/DMUL/DADD/DSETP), MUFU, TMA (UTMALDG), mbarrier (SYNCS), local memory (LDL/STL)"
awk '
/Function : /{ if (name != "") print_row(); name=$3; n=dfma=dmul=dadd=dsetp=mufu=tma=syncs=ldl=stl=0; next }
/^ *\/\*[0-9a-f]+\*\/ /{ n++
  if ($0 ~ / DFMA/) dfma++; if ($0 ~ / DMUL/) dmul++; if ($0 ~ / DADD/) dadd++; if ($0 ~ / DSETP/) dsetp++
  if ($0 ~ / MUFU/) mufu++; if ($0 ~ /UTMALDG/) tma++; if ($0 ~ /SYNCS/) syncs++; if ($0 ~ / LDL/) ldl++; if ($0 ~ / STL/) stl++ }
function print_row() { printf "%-110s inst %6d  DFMA %5d DMUL %5d DADD %5d DSETP %4d  MUFU %3d  UTMALDG %2d SYNCS %2d  LDL %3d STL %3d\n", name, n, dfma, dmul, dadd, dsetp, mufu, tma, syncs, ldl, stl }
END { print_row() }' $TMP | c++filt | sed 's/picles:://g' | sort
echo
echo "# the TMA / mbarrier instructions of k_project_remesh<2> (narrow tiles), verbatim"
awk '/Function : .*k_project_remeshILi2E/{f=1} f&&/Function : /&&!/k_project_remeshILi2E/{f=0} f' $TMP | grep -E "UTMALDG|SYNCS|UBLKCP|FENCE" | sed 's/^ *//' | cut -c1-140
echo
echo "# resource usage (cuobjdump -res-usage)"
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 -E "Function (.*k_advance|.*k_project_remesh|.*k_wind_sample|.*k_seed)" | grep -v "^--" | c++filt | sed 's/picles:://g' | cut -c1-220
rm -f $TMP
