#!/bin/bash
# scaling check (gpurun --gpus 8): bench at N = 4 and 8 (N = 1, 2 come from tools/gpu_mgpu.sh)
mkdir -p gpurun_out
nvidia-smi -L
for N in ${SCALE_NS:-4 8}; do
  echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
done
