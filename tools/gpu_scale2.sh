#!/bin/bash
# bench at N ranks (gpurun --gpus N)
N=${NGPU:-8}
mkdir -p gpurun_out
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
