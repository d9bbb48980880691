#!/bin/bash
# round-1 final evidence: launch lists of the bench command (cfg2, cfg3) and one full ncu capture each of the two
# kernels of the step (k_fused at cfg2, k_pyramid_tiled at cfg3).  Every ncu pass follows a plain run that exited 0.
TAG=${1:-r1g}
mkdir -p gpurun_out
for wl in cfg2 cfg3; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > gpurun_out/plain_$wl.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 40 --csv --log-file gpurun_out/launches_${TAG}_$wl.csv \
      python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > gpurun_out/ncu_list_$wl.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -f -o gpurun_out/prof_${TAG}_fused \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pyramid_tiled -s 30 -c 1 -f -o gpurun_out/prof_${TAG}_pyr \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload cfg3 > gpurun_out/ncu_full2.log 2>&1
tail -n 2 gpurun_out/ncu_full.log; tail -n 2 gpurun_out/ncu_full2.log
