#!/bin/bash
# One gpurun call, what the driver does at round end + the evidence under profiles/: GPU tests, smoke, bench (both arms),
# per-config perf, ncu launch list + full captures of the sweep kernels.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== perf"; timeout 900 python tools/gpu_perf.py c1 c5 c4 c2 c3 > gpurun_out/perf.log 2>&1; echo "perf rc=$?"; grep -A2 "rep1\|rep2\|C5b" gpurun_out/perf.log | grep -v "^--" | cut -c1-400
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c5b > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full sweeps"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_sweep -c 6 -o gpurun_out/prof_sweep -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5b > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full F=12 tma sweep"; PROBE_K=10 PROBE_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_sweep_tma -s 2 -c 2 -o gpurun_out/prof_sweep12_tma -f python tools/probe_sweep.py > gpurun_out/ncu_sweep12.log 2>&1; echo "rc=$?"
ls -la gpurun_out
