#!/bin/bash
# A/B of k_strip builds on one box: profiles/variants/libxptwarp_<tag>.so ("tree" = in-tree, "tiles" = tile kernel)
for tag in "$@"; do
  if [ "$tag" = tree ]; then unset XPTWARP_LIB; timeout 120 python profiles/strip_time.py
  elif [ "$tag" = tiles ]; then unset XPTWARP_LIB; timeout 120 python profiles/strip_time.py tiles
  else XPTWARP_LIB=$PWD/profiles/variants/libxptwarp_$tag.so timeout 120 python profiles/strip_time.py; fi
done 2>&1 | grep -v Warning
