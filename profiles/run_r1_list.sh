#!/bin/bash
mkdir -p gpurun_out
for wl in cfg2 cfg3; do
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > gpurun_out/plain_$wl.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 40 --csv --log-file gpurun_out/launches_$wl.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > gpurun_out/ncu_list_$wl.log 2>&1
done
