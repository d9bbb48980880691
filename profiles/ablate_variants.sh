#!/bin/bash
# Timing ablations of k_fused (XPT_EXP knobs; variant libraries are built HERE before gpurun, results of the
# variants are wrong by construction -- only the kernel time is read).
#   build:  bash profiles/ablate_variants.sh build      run (GPU box): bash profiles/ablate_variants.sh run
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p profiles/variants
  for v in 1 2 4 8 6 12 14; do
    make -s -C xpt-mde-2021_b200/csrc OUT=../../profiles/variants/libxptwarp_exp$v.so EXTRA=-DXPT_EXP=$v
  done
  exit 0
fi
mkdir -p gpurun_out
for v in 1 2 4 8 6 12 14; do
  echo "== XPT_EXP=$v"
  XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_exp$v.so timeout 300 python profiles/phase_split.py ${2:-cfg2} 2>&1 | head -1
done | tee gpurun_out/ablate_variants.txt
