#!/bin/bash
# GPU-box recipe used during round 1 (run under gpurun from the repo root).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --steps 50 --workload cfg3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
