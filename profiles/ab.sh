#!/bin/bash
# A/B of kernel variants on one box: profiles/variants/libxptwarp_<tag>.so vs the in-tree library.
# usage (GPU box): bash profiles/ab.sh tagA tagB ...   ("tree" = the in-tree build)
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for rep in 1 2; do
for tag in "$@"; do
  for wl in cfg2 cfg3; do
    if [ "$tag" = tree ]; then unset XPTWARP_LIB; else export XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_$tag.so; fi
    echo "$tag $wl $(timeout 300 python profiles/phase_split.py $wl 2>&1 | head -2 | tr '\n' '|')" | tee -a gpurun_out/ab.txt
  done
done
done
