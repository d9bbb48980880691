#!/bin/bash
# the final tree under torchrun exactly as the driver launches it (2 ranks): bench line incl. the two-in-flight e2e leg, reference arm
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 200 --warmup 20 2>gpurun_out/n2check.err | tail -1 | tee gpurun_out/bench_n2_final.json | cut -c1-300
python -c "import json; d=json.load(open('gpurun_out/bench_n2_final.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['sync_value'], d['warmup'], d['config']['collectives'][:50])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 2 --impl reference --steps 3 --warmup 1 2>>gpurun_out/n2check.err | tail -1 | cut -c1-200
tail -n 4 gpurun_out/n2check.err
