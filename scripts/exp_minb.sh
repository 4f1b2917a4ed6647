#!/bin/bash
# Rebuild the library on the GPU box with different register caps for the one-warp-per-transform kernels and time them.
set -u
mkdir -p gpurun_out/exp
for MB in 1 5 6; do
  make -B -C differentiable-mel-spectrogram_b200/csrc EXTRA=-DDMEL_MINB_SMALL=$MB > /dev/null 2>&1
  echo "== MINB=$MB"
  python scripts/tune.py ${@:-audio_mnist} 2>&1 | grep -E '"ng": 4, "sf": (2|4),'
done 2>&1 | tee gpurun_out/exp/minb.log
make -B -C differentiable-mel-spectrogram_b200/csrc > /dev/null 2>&1
