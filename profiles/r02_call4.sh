#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL tests after the boundary-launch change, strong scaling at N=2, C5 strips at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_gputests3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gputests3.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2_b.json'))
for k in ("value","ms_per_step","gpu_launches","strong","strip_parity"):
    print(k, json.dumps(d.get(k)))
PY
python profiles/bench_strips.py --config C5 > gpurun_out/r02_strips_c5_n1.json 2>gpurun_out/r02_strips.err; cat gpurun_out/r02_strips_c5_n1.json | cut -c1-700
$TR profiles/bench_strips.py --config C5 > gpurun_out/r02_strips_c5_n2.json 2>>gpurun_out/r02_strips.err; cat gpurun_out/r02_strips_c5_n2.json | cut -c1-900
$TR profiles/bench_strips.py --config C5 --balance > gpurun_out/r02_strips_c5_n2b.json 2>>gpurun_out/r02_strips.err; cat gpurun_out/r02_strips_c5_n2b.json | cut -c1-900
