#!/bin/bash
# exploratory 1-GPU run: parity tests, then the C5 bench under a few tuning knobs
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in 1e12 2e5 1e5 2e4; do
  echo "== bench CL_WORK=$w"
  BESS_B200_CL_WORK=$w timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c5b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'fits/s', round(d['value']), {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
echo "== perf"; timeout 900 python tools/gpu_perf.py ${PERF_CFGS:-c4 c2 c3} > gpurun_out/perf.log 2>&1; echo "perf rc=$?"; grep -v "true support" gpurun_out/perf.log | grep "rep1\|prof_ms" | tail -12
