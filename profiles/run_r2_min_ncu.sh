#!/bin/bash
# one full ncu capture of k_min_strip<GRAD, PAIR> (MoA pair, config-2 size)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_min_strip -s 4 -c 1 -f -o gpurun_out/prof_min_pair \
    python profiles/minloss_one.py PAIR moa > gpurun_out/ncu_min_pair.log 2>&1
tail -1 gpurun_out/ncu_min_pair.log
