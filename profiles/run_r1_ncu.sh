#!/bin/bash
# ncu recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture of the top kernel.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -o gpurun_out/prof_fused \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
