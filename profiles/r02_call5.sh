#!/bin/bash
# round 2, GPU call 5 (8 GPUs): NCCL tests on 3-4 GPUs, bench at N=8 (weak + strong + parity), C5 strong scaling at N=8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_gputests_multi8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests_multi8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n8.json'))
for k in ("value","ms_per_step","gpu_launches","strong","strip_parity","e2e","e2e_store","e2e_wind_mesh"):
    print(k, json.dumps(d.get(k)))
PY
$TR profiles/bench_strips.py --config C5 > gpurun_out/r02_strips_c5_n8.json 2>gpurun_out/r02_strips8.err; cat gpurun_out/r02_strips_c5_n8.json | cut -c1-1400
$TR profiles/bench_strips.py --config C5 --balance > gpurun_out/r02_strips_c5_n8b.json 2>>gpurun_out/r02_strips8.err; cat gpurun_out/r02_strips_c5_n8b.json | cut -c1-1400
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
$TR4 profiles/bench_strips.py --config C5 --balance > gpurun_out/r02_strips_c5_n4b.json 2>>gpurun_out/r02_strips8.err; cat gpurun_out/r02_strips_c5_n4b.json | cut -c1-900
