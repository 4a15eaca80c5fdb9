#!/bin/bash
# round 2, GPU call 8 (1 GPU): the ncu evidence of the committed build — full captures of the advance and gather kernels
# (step 4), of the wind sampler and of the AutoTsit5 instantiation; launch list of a bench run (taken after the same
# command has exited 0 without ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
python profiles/prof_step.py 4096 5 > gpurun_out/r02_prof5.log 2>&1; echo "prof rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_advance|k_project_remesh' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_step4 python profiles/prof_step.py 4096 5 > gpurun_out/r02_ncu_step4.log 2>&1; echo "ncu step4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_wind' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r02_wind python profiles/prof_wind.py > gpurun_out/r02_ncu_wind.log 2>&1; echo "ncu wind rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_advance' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_auto_step4 python profiles/prof_step_auto.py 4096 5 > gpurun_out/r02_ncu_auto.log 2>&1; echo "ncu auto rc=$?"
