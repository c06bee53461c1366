#!/bin/bash
# exploratory 1-GPU run: parity tests, then sweep probes / the C5 bench / all-config perf under both sweep kernels
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for sw in stream tma; do
  echo "== sweep=$sw probe"; BESS_B200_SWEEP=$sw python tools/probe_sweep.py
  echo "== sweep=$sw bench"
  BESS_B200_SWEEP=$sw timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'fits/s', round(d['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
print('c5b', d['c5b_no_screening']['ms_per_call'], 'probe', d['roofline']['p500k_pdas_sweep']['achieved'])"
  echo "== sweep=$sw perf"; BESS_B200_SWEEP=$sw timeout 900 python tools/gpu_perf.py ${PERF_CFGS:-c4 c2 c3} > gpurun_out/perf_$sw.log 2>&1; echo "perf rc=$?"; grep -A2 "rep1" gpurun_out/perf_$sw.log | grep -v "^--"
done
