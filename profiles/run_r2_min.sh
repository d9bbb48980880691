#!/bin/bash
# round 2: the min-over-sources strip kernel -- parity (new vs old kernel, pair launch, oracle, goldens), then device time
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "min_ or stereo_total or golden" 2>&1 | tail -15 | tee gpurun_out/pytest_min.txt
timeout 300 python profiles/minloss_bench.py 2>&1 | tail -8 | tee gpurun_out/minloss_strip.txt
timeout 300 python profiles/loss_sets.py 2>&1 | tail -12 | tee gpurun_out/loss_sets.txt
