#!/bin/bash
# round 2, GPU call 23 (2 GPUs, the round's last GPU seconds): the interior advance leaves four SMs to the exchange,
# communication stream at its normal priority — bench legs at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n2_c.json 2> gpurun_out/r02_bench_n2_c.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2_c.json'))
print("value", d["value"], d["ms_per_step"])
for k in ("strong","strip_parity","strong_c5"):
    v=d.get(k) or {}
    print(k, {q: v.get(q) for q in ("efficiency","ms_per_step","ms_per_step_1gpu","result","rows_per_rank","ms_per_step_per_rank")})
PY
