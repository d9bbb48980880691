#!/bin/bash
# round-2 multi-GPU evidence (N from $1): the 2-GPU NCCL parity test, then the bench exactly as the driver launches it
# for cfg2 (weak), cfg4 (B=64 global, strong) and cfg5 with the full backward (B=128 global, strong)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "two_gpu" 2>&1 | tail -3 | tee gpurun_out/pytest_two_gpu_n${N}.txt
run() {  # name port args...
  local name=$1 port=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N "$@" 2>gpurun_out/bench_n${N}_$name.err | tail -1 | tee gpurun_out/bench_n${N}_$name.json
}
run cfg2 29511 --steps 200 --warmup 20
run cfg4 29512 --steps 200 --warmup 20 --workload cfg4
run cfg5_srcgrad 29513 --steps 20 --warmup 5 --workload cfg5 --source-grad
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_n${N}_ref.err | tail -1 | tee gpurun_out/bench_n${N}_reference.json
for e in gpurun_out/bench_n${N}_*.err; do tail -n 3 "$e"; done
