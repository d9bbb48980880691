#!/bin/bash
# A/B of the fused kernel's grid order: XPT_GRID_TILE_MAJOR = 0 (snippet-major), 1 (tile-major), 2 (by grid size: the default)
mkdir -p gpurun_out
: > gpurun_out/ab_tm.txt
for rep in 1 2; do for v in 0 1 2; do
  XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_tm$v.so timeout 400 python profiles/ab_time.py cfg2 cfg3 2>&1 | tail -2 | tee -a gpurun_out/ab_tm.txt
done; done
