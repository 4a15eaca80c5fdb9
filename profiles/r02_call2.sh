#!/bin/bash
# round 2, GPU call 2: GPU test suite, the work-queue advance on every config shape, DP5 / wind-sampler variants, bench line, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gputests.log
python profiles/bench_configs.py --steps 5 --warmup 3 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err; echo "configs rc=$?"
cat gpurun_out/r02_configs.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:40], '%.3e'%d['particle_steps_per_s'], 'adv %.2f prj %.2f'%(d['ms_advance'],d['ms_project_remesh']), d['max_attempts'])"
PROF=prof_step_dp5.py bash profiles/variants.sh time base dp5_ct base
for v in base wind_row4; do echo "wind $v"; PICLES_B200_LIB=$PWD/_exp/lib_$v.so python profiles/prof_wind.py; done
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r02_bench_n1_a.json
ncu --set full --clock-control none --import-source on -k regex:'k_advance|k_project_remesh' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_step4 python profiles/prof_step.py 4096 5 > gpurun_out/r02_ncu_step4.log 2>&1; echo "ncu rc=$?"
