#!/bin/bash
# round 2, GPU call 15 (8 GPUs): bench.py at N=8 as the driver launches it — weak value, strong (fixed 4096^2 box),
# strip_parity over NCCL, strong_c5 (tripolar + land grid in 8 strips by measured cost)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err ); echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n8.json'))
for k in ("value","ms_per_step","gpu_launches","strong","strip_parity","strong_c5","e2e","e2e_store","e2e_wind_mesh"):
    print(k, json.dumps(d.get(k))[:1200])
PY
tail -3 gpurun_out/r02_bench_n8.err
