#!/bin/bash
# round 2, GPU call 21 (1 GPU): last check of the final build — smoke(), the strip tests on one GPU, a short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_oned.py -m gpu -q -x -k "not headline and not medium and not large" ) > gpurun_out/r02_gputests21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputests21.log | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_last.json 2> gpurun_out/r02_bench_n1_last.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_last.json')); print('value %.4e ms %.3f frac %.4f e2e %.4e auto %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['default_solver_variant']['ms_advance']))"
