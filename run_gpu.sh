python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 30 -c 1 -o gpurun_out/prof_r1b python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
