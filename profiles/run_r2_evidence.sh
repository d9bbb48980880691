#!/bin/bash
# round-2 evidence on one B200: full parity suite, bench lines of every BASELINE config at N=1, pyramid A/B, reference arm,
# then the ncu recipe (plain run first, launch list, one --set full capture per kernel and workload)
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $O/pytest_gpu.txt
b() { name=$1; shift; timeout 600 python bench.py "$@" 2>$O/bench_$name.err | tail -1 | tee $O/bench_$name.json; }
b cfg2 --steps 200 --warmup 20
b cfg3 --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline
b cfg4 --steps 100 --warmup 10 --workload cfg4 --no-cpu-baseline
b cfg5 --steps 10 --warmup 3 --workload cfg5 --no-cpu-baseline
b cfg5_srcgrad --steps 10 --warmup 3 --workload cfg5 --source-grad --no-cpu-baseline
XPT_PYRAMID=tiled b cfg2_pyr_tiled --steps 200 --warmup 20 --no-cpu-baseline
XPT_PYRAMID=tiled b cfg3_pyr_tiled --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline
b reference --impl reference --steps 3 --warmup 1
for wl in cfg2 cfg3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --kill 1 --csv --log-file $O/launches_$wl.csv \
      python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > $O/ncu_list_$wl.log 2>&1
done
for wl in cfg2 cfg3 cfg5; do
  timeout 900 ncu --set full --clock-control none --import-source on --kill 1 -k regex:k_fused -s 12 -c 1 -f -o $O/prof_fused_$wl \
      python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > $O/ncu_fused_$wl.log 2>&1
  tail -1 $O/ncu_fused_$wl.log
done
for wl in cfg2 cfg3; do
  timeout 900 ncu --set full --clock-control none --import-source on --kill 1 -k regex:k_pyramid_tma -s 12 -c 1 -f -o $O/prof_pyr_$wl \
      python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline --workload $wl > $O/ncu_pyr_$wl.log 2>&1
  tail -1 $O/ncu_pyr_$wl.log
done
ls -la $O | head -40
