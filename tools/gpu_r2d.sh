#!/bin/bash
mkdir -p gpurun_out
grep -E "MemTotal|MemAvailable" /proc/meminfo; nproc
echo "== bench ref"; timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo done
