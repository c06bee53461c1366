#!/bin/bash
# exploratory 1-GPU run: parity tests, chain_fit phase timers, bench, all-config perf
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== phases"; timeout 600 python tools/gpu_phase.py ${PHASE_CFGS:-c2} 2>&1 | tail -14
echo "== probe"; python tools/probe_sweep.py
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'fits/s', round(d['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
print('c5b', d['c5b_no_screening']['ms_per_call'], 'probe', d['roofline']['p500k_pdas_sweep']['achieved'])"
echo "== perf"; timeout 900 python tools/gpu_perf.py ${PERF_CFGS:-c4 c2 c3} > gpurun_out/perf.log 2>&1; echo "perf rc=$?"; grep -A2 "rep1" gpurun_out/perf.log | grep -v "^--"
