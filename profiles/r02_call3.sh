#!/bin/bash
# round 2, GPU call 3 (2 GPUs): whole GPU suite incl. the NCCL tests, bench at N=2 with the strong / parity legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_gputests2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_gputests2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2_a.json'))
for k in ("value","ms_per_step","gpu_launches","strong","strip_parity","e2e","e2e_store"):
    print(k, json.dumps(d.get(k)))
PY
tail -5 gpurun_out/r02_bench_n2_a.err
