#!/bin/bash
# Round-2 evidence in one gpurun call: per-config perf, ncu --set full of the GLM / Cox batched sweeps (C2, C4), of the
# screening sweep and of the resident-path kernel, ncu launch list of a C5 step.
mkdir -p gpurun_out
echo "== perf all configs"; timeout 1200 python tools/gpu_perf.py c1 c5 c4 c2 c3 > gpurun_out/perf.log 2>&1; echo "rc=$?"; grep -A1 "rep1" gpurun_out/perf.log | cut -c1-260
echo "== ncu GLM sweep (C2: dual_sweep_tma<12,MODE_DH>)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_sweep_tma -s 6 -c 1 -o gpurun_out/prof_sweep_c2 -f python tools/gpu_perf.py c2 > gpurun_out/ncu_c2.log 2>&1; echo "rc=$?"
echo "== ncu Cox sweep (C4: dual_sweep_tma<6,MODE_COX>)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_sweep_tma -s 6 -c 1 -o gpurun_out/prof_sweep_c4 -f python tools/gpu_perf.py c4 > gpurun_out/ncu_c4.log 2>&1; echo "rc=$?"
echo "== ncu launch list C5"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5b > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: screening sweep + resident kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dual_sweep_kernel|lm_path_kernel" -s 2 -c 4 -o gpurun_out/prof_c5_main -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c5b > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep ncu-rep
