#!/bin/bash
# one full ncu capture (with source) of the k_fused launch of the default bench command; $1 = tag, $2 = workload
TAG=${1:-x}; WL=${2:-cfg2}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $WL > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $WL > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
