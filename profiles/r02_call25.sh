#!/bin/bash
# round 2, GPU call 25 (1 GPU, the round's last GPU seconds): dp5_blowup in two strips and through a checkpoint
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dp5_blowup" > gpurun_out/r02_gputests25.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_gputests25.log | cut -c1-300
