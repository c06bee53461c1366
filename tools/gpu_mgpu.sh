#!/bin/bash
# multi-GPU check (gpurun --gpus N): NCCL-sharded parity worker + bench at N
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== sharded parity worker"; NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > gpurun_out/mgpu_worker.log 2>&1; echo "worker rc=$?"; tail -25 gpurun_out/mgpu_worker.log
echo "== nccl probe"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/nccl_probe.py 2>&1 | grep -v "^\*\|OMP_NUM"; nvidia-smi topo -m 2>&1 | head -12
echo "== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
