#!/bin/bash
# persistent double-buffered TMA pyramid (default) against the grid form (XPT_PYRAMID=tma_grid) and the LDG/STS tile kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "golden or pieces or ragged or edge or config3 or full_size or config5 or host" 2>&1 | tail -3
: > gpurun_out/pyr_ab.txt
for mode in default tma_grid tiled; do
  for wl in cfg2 cfg3; do
    XPT_PYRAMID=$mode timeout 300 python bench.py --steps 200 --warmup 20 --workload $wl --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode $wl value %.3f step %.1f us pyramid %.2f us (%.1f %% of peak)' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['roofline']['secondary']['kernel_ms'], 100*d['roofline']['secondary']['frac']))" | tee -a gpurun_out/pyr_ab.txt
  done
done
