#!/bin/bash
# quick GPU check + one full ncu capture of k_fused (tag from $1)
TAG=${1:-x}
mkdir -p gpurun_out
timeout 120 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg2.json
timeout 300 python bench.py --steps 50 --workload cfg3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg3.json
timeout 300 python profiles/phase_split.py cfg2 2>&1 | tee gpurun_out/phase_cfg2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
