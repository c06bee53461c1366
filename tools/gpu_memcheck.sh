#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: solver / Gram probes, wide groups, rank-deficient fallback, fused
# normalisation, large-support solvers (DSMEM reduce, staged Cholesky)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
CS="compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20"
echo "== solve probe"; IMPLS=0 timeout 600 $CS python tools/gpu_solve_probe.py 17 100 240 > gpurun_out/mc_solve.log 2>&1; echo "rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/mc_solve.log; tail -2 gpurun_out/mc_solve.log
echo "== gram probe"; GRAM_CASES=113x203,250x64,37x130 GRAM_IMPLS=2 timeout 600 $CS python tools/gpu_gram_probe.py > gpurun_out/mc_gram.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/mc_gram.log
echo "== pytest subset"; timeout 1500 $CS python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide_lm_seq_gic or wide_cox_seq_gic or dupsig_lm_seq_gic or dupsig_logit_seq_gic or (large_support and gaussian-700-900-250) or (large_support and binomial-1500-800-150) or lm_seq_cv" > gpurun_out/mc_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/mc_pytest.log
