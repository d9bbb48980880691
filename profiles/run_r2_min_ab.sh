#!/bin/bash
# device time of the min / combined losses: k_min_strip against the round-1 kernel on the same inputs
mkdir -p gpurun_out
timeout 300 python profiles/minloss_bench.py 2>&1 | tail -8 | tee gpurun_out/minloss_strip.txt
timeout 300 python profiles/minloss_bench.py tiles 2>&1 | tail -8 | tee gpurun_out/minloss_tiles.txt
