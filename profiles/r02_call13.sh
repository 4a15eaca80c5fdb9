#!/bin/bash
# round 2, GPU call 13 (2 GPUs): NCCL tests, bench at N=2 with the strong / strip_parity / strong_c5 legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_gputests_multi13.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputests_multi13.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ); echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
for k in ("value","ms_per_step","gpu_launches","strong","strip_parity","strong_c5","e2e","e2e_store"):
    print(k, json.dumps(d.get(k))[:900])
PY
tail -5 gpurun_out/r02_bench_n2.err
