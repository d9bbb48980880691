python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 600 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
python bench.py --steps 50 --workload cfg3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg3.json; cat gpurun_out/bench_cfg3.json
python bench.py --steps 100 --unfused --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg2_unfused.json; cat gpurun_out/bench_cfg2_unfused.json
python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_photo -s 30 -c 2 -o gpurun_out/prof_r1a python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
