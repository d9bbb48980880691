#!/bin/bash
# round 2, final tree: whole GPU suite, smoke, bench lines (cfg2 default incl. CPU baseline, cfg3, cfg5), k_strip A/B,
# one-launch-per-eye A/B of the stereo LOSS_RIGID_T1 step
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.txt
cp gpurun_out/parity_margins.json $O/parity_margins.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
b() { name=$1; shift; timeout 600 python bench.py "$@" 2>$O/bench_$name.err | tail -1 | tee $O/bench_$name.json | cut -c1-200; }
b cfg2 --steps 200 --warmup 20
b cfg3 --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline
b cfg5 --steps 10 --warmup 3 --workload cfg5 --no-cpu-baseline
b cfg2_strip --steps 200 --warmup 20 --strip --no-cpu-baseline
b cfg3_strip --steps 50 --warmup 10 --workload cfg3 --strip --no-cpu-baseline
XPT_LS_ONLY=LOSS_RIGID_T1 timeout 200 python profiles/loss_sets.py 2>&1 | tail -1 | tee $O/t1_same_launch.txt
XPT_EYES=separate XPT_LS_ONLY=LOSS_RIGID_T1 timeout 200 python profiles/loss_sets.py 2>&1 | tail -1 | tee $O/t1_separate.txt
