#!/bin/bash
# one full ncu capture of k_min_strip<GRAD> per method (MoA, config-2 size)
mkdir -p gpurun_out
for m in SSIM L1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_min_strip -s 4 -c 1 -f -o gpurun_out/prof_min_$m \
    python profiles/minloss_one.py $m moa > gpurun_out/ncu_min_$m.log 2>&1
  tail -1 gpurun_out/ncu_min_$m.log
done
ls -la gpurun_out/*.ncu-rep
