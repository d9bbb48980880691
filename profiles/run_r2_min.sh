#!/bin/bash
# round 2: the min-over-sources / combined strip kernel -- parity (new vs old kernel, pair launch, oracle, goldens), device time,
# one full ncu capture of the pair launch
# (per-kernel launch list of one step: XPT_LS_ONLY=LOSS_RIGID_MOA_WST XPT_LS_STEPS=1 ncu --metrics gpu__time_duration.sum ... python profiles/loss_sets.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "min_ or stereo_total or golden" 2>&1 | tail -5 | tee gpurun_out/pytest_min.txt
timeout 300 python profiles/minloss_bench.py 2>&1 | tail -8 | tee gpurun_out/minloss_strip.txt
timeout 300 python profiles/loss_sets.py 2>&1 | tail -12 | tee gpurun_out/loss_sets.txt
timeout 600 ncu --set full --clock-control none --import-source on --kill 1 -k regex:k_min_strip -s 4 -c 1 -f -o gpurun_out/prof_min_pair_final \
    python profiles/minloss_one.py PAIR moa > gpurun_out/ncu_min_pair.log 2>&1
tail -1 gpurun_out/ncu_min_pair.log
