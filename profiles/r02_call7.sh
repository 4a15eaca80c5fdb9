#!/bin/bash
# round 2, GPU call 7 (1 GPU): tripolar gather after the index-math change; ncu of the C5 gather
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python profiles/bench_configs.py --steps 5 --warmup 3 --only "tripolar" > gpurun_out/r02_configs_tri.jsonl 2> gpurun_out/r02_configs_tri.err; echo "configs rc=$?"
python -c "
import sys,json
for l in open('gpurun_out/r02_configs_tri.jsonl'):
    d=json.loads(l); print(d['config'][:40], '%.3e'%d['particle_steps_per_s'], 'adv %.2f prj %.3f'%(d['ms_advance'],d['ms_project_remesh']), d['max_attempts'], d['reach'])"
ncu --set full --clock-control none --import-source on -k regex:'k_project_remesh' --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_c5_gather python profiles/bench_configs.py --only "C5" --steps 3 --warmup 3 > gpurun_out/r02_ncu_c5.log 2>&1; echo "ncu rc=$?"
