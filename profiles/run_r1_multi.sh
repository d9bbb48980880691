#!/bin/bash
# multi-GPU bench exactly as the driver launches it (N from $1); every run under its own timeout
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 200 --warmup 20 2>gpurun_out/bench_n${N}.err | tail -1 | tee gpurun_out/bench_n${N}.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 50 --warmup 10 --workload cfg3 2>>gpurun_out/bench_n${N}.err | tail -1 | tee gpurun_out/bench_n${N}_cfg3.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_n${N}.err | tail -1
tail -3 gpurun_out/bench_n${N}.err
