#!/bin/bash
# ncu captures: (1) F=12 dual sweep at p=500k, (2) chain_fit / topk / finish kernels of a C5 step, (3) launch list of a C5 step
mkdir -p gpurun_out
echo "== probe (no profiler)"; python tools/probe_sweep.py
echo "== ncu sweep F=12"; PROBE_K=10 PROBE_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_sweep_kernel -s 4 -c 2 -o gpurun_out/prof_sweep12 -f python tools/probe_sweep.py > gpurun_out/ncu_sweep12.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_sweep12.log
echo "== ncu chain/topk/finish"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_fit|topk_slices|finish_kernel" -s 60 -c 6 -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5b > gpurun_out/ncu_small.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_small.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c5b > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -30
