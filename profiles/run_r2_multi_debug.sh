#!/bin/bash
# short 2-GPU check: the NCCL parity test with its full traceback, then the cfg2 bench line as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -k "two_gpu" 2>&1 | tail -60 | cut -c1-300 | tee gpurun_out/pytest_two_gpu_n${N}.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/bench_n${N}_cfg2.err | tail -1 | tee gpurun_out/bench_n${N}_cfg2.json
tail -5 gpurun_out/bench_n${N}_cfg2.err | cut -c1-300
