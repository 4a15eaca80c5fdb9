#!/bin/bash
# round 2, GPU call 24 (1 GPU, the round's last GPU minute): the dp5_blowup scenario (trial steps whose stages overflow,
# DtNaN) through the CUDA path against the oracle, then smoke() of the rebuilt library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dp5_blowup or nan_defaults" > gpurun_out/r02_gputests24.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_gputests24.log | cut -c1-300
timeout 15 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
