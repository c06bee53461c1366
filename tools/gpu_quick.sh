#!/bin/bash
# quick GPU check: parity tests + bench + per-config perf
mkdir -p gpurun_out
echo "== pytest gpu"; timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps ${BENCH_STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== perf"; timeout 900 python tools/gpu_perf.py ${PERF_CFGS:-c1 c5} > gpurun_out/perf.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf.log
