#!/bin/bash
# round 2: parity tests on the in-tree build, then A/B of the XPT_YPIPE variants (profiles/variants/libxptwarp_yp<N>.so)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
: > gpurun_out/ab_yp.txt
for rep in 1 2; do for v in 0 1 2 3; do
  XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_yp$v.so timeout 200 python profiles/ab_time.py cfg2 cfg3 2>&1 | tail -2 | tee -a gpurun_out/ab_yp.txt
done; done
