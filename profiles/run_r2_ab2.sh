#!/bin/bash
# A/B of grid order (XPT_GRID_TILE_MAJOR) on cfg2, cfg3, cfg5; also validates the in-tree build's pyramid change
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "golden or config3 or full_size or ragged" 2>&1 | tail -3
: > gpurun_out/ab_tm.txt
for rep in 1 2; do for v in 0 1; do
  XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_tm$v.so timeout 400 python profiles/ab_time.py cfg2 cfg3 cfg5 2>&1 | tail -3 | tee -a gpurun_out/ab_tm.txt
done; done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 value', d['value'], 'pyr', d['roofline']['secondary']['kernel_ms'])"
timeout 300 python bench.py --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 value', d['value'], 'pyr', d['roofline']['secondary']['kernel_ms'])"
