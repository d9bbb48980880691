#!/bin/bash
# quick GPU check of a kernel change: parity tests, cfg2 + cfg3 bench lines, ablation
mkdir -p gpurun_out
timeout 120 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg2.json
timeout 300 python bench.py --steps 50 --workload cfg3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg3.json
timeout 300 python profiles/phase_split.py cfg2 2>&1 | tee gpurun_out/phase_cfg2.txt
