#!/bin/bash
# per-kernel device time of the object-by-object rows (MoA / MonoDepth2 / Combined / flow) at config-2 frame size
mkdir -p gpurun_out
timeout 300 python profiles/loss_sets.py > gpurun_out/loss_sets.txt 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_loss_sets.csv \
    python profiles/loss_sets.py > gpurun_out/ncu_loss_sets.log 2>&1
tail -n 7 gpurun_out/loss_sets.txt
