#!/bin/bash
# round 2, GPU call 1: the prepared switches of the advance kernel side by side, and the
# SIMT divergence of k_advance on the config shapes with uneven attempt counts (C3, C4, C5).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_call1_smi.txt
bash profiles/variants.sh time base stage_switch share_rcp reg_sums hoist_sdir all4 base > gpurun_out/r02_variants.txt 2>&1
cat gpurun_out/r02_variants.txt
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for c in "C3 growing/decaying winds 2048x2048, on" "C4 tripolar" "C5 tripolar"; do
  tag=$(echo "$c" | cut -c1-2)
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_advance --launch-skip 3 --launch-count 2 --csv \
      --log-file gpurun_out/r02_div_$tag.csv python profiles/bench_configs.py --only "$c" --steps 2 --warmup 3 > gpurun_out/r02_div_$tag.log 2>&1
  echo "== $tag rc=$?"; tail -3 gpurun_out/r02_div_$tag.csv
done
