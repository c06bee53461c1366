#!/bin/bash
# multi-GPU check (gpurun --gpus N): NCCL-sharded parity worker + bench at N
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== sharded parity worker"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; echo "worker rc=$?"; tail -14 gpurun_out/mgpu_worker.log
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
