#!/bin/bash
# round 2, first look at the resident path: LM parity tests, host-phase split, short bench
mkdir -p gpurun_out
echo "== pytest gpu (lm first)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lm or gauss or golden or c5 or c1 or topk or pipelined" > gpurun_out/pytest_lm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_lm.log
echo "== hostphase"; timeout 300 python tools/gpu_hostphase.py > gpurun_out/hostphase.log 2>&1; echo "rc=$?"; cat gpurun_out/hostphase.log | tail -8
echo "== pytest gpu all"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
