#!/bin/bash
# e2e leg with 2 / 3 / 4 host-buffer steps in flight
mkdir -p gpurun_out
for d in 2 3 4; do
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-depth $d 2>gpurun_out/e2e.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 value', d['value'], 'e2e', d['e2e']['in_flight'], d['e2e']['value'], d['e2e']['sync_value'])"
done
timeout 600 python bench.py --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline --e2e-depth 3 2>>gpurun_out/e2e.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 value', d['value'], 'e2e', d['e2e']['in_flight'], d['e2e']['value'], d['e2e']['sync_value'])"
tail -3 gpurun_out/e2e.err
