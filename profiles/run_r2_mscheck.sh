#!/bin/bash
# the min / combined strip kernel with device-side bounds assertions (XPT_MS_CHECK build) through its parity tests
mkdir -p gpurun_out
XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_mscheck.so timeout 600 python -m pytest tests -x -q -m gpu -k "min_ or stereo_total or flow_total or combined" 2>&1 | tail -4 | tee gpurun_out/mscheck.txt
