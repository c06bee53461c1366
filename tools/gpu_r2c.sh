#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== hostphase"; timeout 300 python tools/gpu_hostphase.py > gpurun_out/hostphase.log 2>&1; echo "rc=$?"; cat gpurun_out/hostphase.log | tail -8
