#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "host" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/e2e.err | tail -1 | tee gpurun_out/bench_e2e_cfg2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 value', d['value'], 'e2e', d['e2e'])"
timeout 600 python bench.py --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline 2>>gpurun_out/e2e.err | tail -1 | tee gpurun_out/bench_e2e_cfg3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 value', d['value'], 'e2e', d['e2e'])"
tail -3 gpurun_out/e2e.err
