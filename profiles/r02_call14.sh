#!/bin/bash
# round 2, GPU call 14 (1 GPU): the 1-D model on the device, the whole GPU suite, the bench line of the committed
# build, and its ncu evidence (launch list of the bench, full captures of advance / gather / wind sampler / AutoTsit5)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_oned.py -m gpu -q -x ) > gpurun_out/r02_gputests_oned.log 2>&1; echo "pytest 1-D rc=$?"; tail -15 gpurun_out/r02_gputests_oned.log | cut -c1-220
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_gputests14.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests14.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print("value %.4e ms %.3f frac %.4f e2e %.4e (%.3f ms) store %.4e (%.3f ms) mesh %.4e launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e_store"]["value"],d["e2e_store"]["ms_per_step"],d["e2e_wind_mesh"]["value"],d["gpu_launches"]))
print("  hbm", [(r["kernel"], round(r["frac"],3), round(r["ms_per_launch"],4)) for r in d["roofline_hbm"]], d.get("cpu_baseline",{}).get("value"), d["default_solver_variant"]["ms_advance"], d["clocks"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_advance|k_project_remesh' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_step4 python profiles/prof_step.py 4096 5 > gpurun_out/r02_ncu_step4.log 2>&1; echo "ncu step4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_wind' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r02_wind python profiles/prof_wind.py > gpurun_out/r02_ncu_wind.log 2>&1; echo "ncu wind rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_advance' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_auto_step4 python profiles/prof_step_auto.py 4096 5 > gpurun_out/r02_ncu_auto.log 2>&1; echo "ncu auto rc=$?"
ls -la gpurun_out/*.ncu-rep
