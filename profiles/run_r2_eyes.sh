#!/bin/bash
# stereo loss sets with one launch per eye / direction (default) against the torch.cat + shared-launch path (XPT_EYES=cat)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "stereo or call_surface or cuda_graph" 2>&1 | tail -3
for e in separate cat; do
  XPT_EYES=$e XPT_LS_ONLY=LOSS_RIGID_T1 timeout 200 python profiles/loss_sets.py 2>&1 | tail -1 | sed "s/^/XPT_EYES=$e /" | tee -a gpurun_out/eyes.txt
  XPT_EYES=$e XPT_LS_ONLY=LOSS_RIGID_MOA_WST timeout 200 python profiles/loss_sets.py 2>&1 | tail -1 | sed "s/^/XPT_EYES=$e /" | tee -a gpurun_out/eyes.txt
done
