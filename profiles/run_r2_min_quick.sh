#!/bin/bash
# quick check of a k_min_strip change: its parity tests and device time
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "min_ or stereo_total or flow_total or combined" 2>&1 | tail -3 | tee gpurun_out/pytest_min.txt
timeout 300 python profiles/minloss_bench.py 2>&1 | tail -8 | tee gpurun_out/minloss_strip.txt
timeout 300 python profiles/loss_sets.py 2>&1 | tail -6 | tee gpurun_out/loss_sets.txt
