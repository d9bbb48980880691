#!/bin/bash
# round 1 session 2: tests, bench, ablation, launch list and a full ncu capture (with source) of the current k_fused
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --steps 50 --workload cfg3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_cfg3.json
python profiles/phase_split.py cfg2 2>&1 | tee gpurun_out/phase_cfg2.txt
python profiles/phase_split.py cfg3 2>&1 | tee gpurun_out/phase_cfg3.txt
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -f -o gpurun_out/prof_r1d \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
