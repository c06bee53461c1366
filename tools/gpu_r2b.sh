#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (lm first)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lm or gauss or golden or c5 or c1 or topk or pipelined" > gpurun_out/pytest_lm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_lm.log
echo "== hostphase"; timeout 300 python tools/gpu_hostphase.py > gpurun_out/hostphase.log 2>&1; echo "rc=$?"; cat gpurun_out/hostphase.log | tail -8
