#!/bin/bash
# Round-2 final evidence in one gpurun call: GPU tests, smoke, bench (both arms), per-config perf, building-block probes,
# ncu --set full of one chain_fit launch at config 2, ncu launch list of a C5 step.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/smoke.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== perf all configs"; timeout 900 python tools/gpu_perf.py c1 c5 c4 c2 c3 > gpurun_out/perf.log 2>&1; echo "rc=$?"; grep -A1 "rep1" gpurun_out/perf.log | cut -c1-200
echo "== probes"; ([ -x tools/micro/fp64_rate ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_rate.cu -o tools/micro/fp64_rate; ./tools/micro/fp64_rate; timeout 200 python tools/gpu_solve_probe.py; timeout 200 python tools/gpu_gram_probe.py; python tools/gpu_phase.py c2) > gpurun_out/probes.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/probes.log
echo "== ncu chain_fit C2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_fit_kernel -s 40 -c 1 -o gpurun_out/prof_chain_c2_r2 -f python tools/gpu_perf.py c2 > gpurun_out/ncu_chain_c2_r2.log 2>&1; echo "rc=$?"
echo "== ncu launch list C5"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5b --no-c2 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep ncu-rep
