#!/bin/bash
# round 2, GPU call 16 (2 GPUs): boundary zones + early all-reduce of their reach in picles_step_strip — the NCCL strip
# tests (overlapped path: fast_box with a widening exchange, tripolar_tall), single-GPU strip tests, bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_gputests16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gputests16.log | cut -c1-300
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err ); echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2_b.json'))
print("value", d["value"], d["ms_per_step"])
for k in ("strong","strip_parity","strong_c5"):
    v=d.get(k) or {}
    print(k, {q: v.get(q) for q in ("efficiency","ms_per_step","ms_per_step_1gpu","result","rows_per_rank","ms_per_step_per_rank","halo_rows_exchanged")})
PY
tail -3 gpurun_out/r02_bench_n2_b.err
