#!/bin/bash
# Experiment: dx-producing backward at 5 CTAs/SM (96 registers + 8-frame tiles so shared memory fits five)
export TUNE_DEFAULT=1
echo "== baseline"; python scripts/tune.py audio_mnist sweep_512 2>&1 | tail -2
make -B -C differentiable-mel-spectrogram_b200/csrc EXTRA=-DDMEL_FULL_BWD_CAP > /dev/null 2>&1
echo "== capped, default tile"; python scripts/tune.py audio_mnist sweep_512 2>&1 | tail -2
echo "== capped, SF=2"; DMEL_NG=4 DMEL_SF=2 python scripts/tune.py audio_mnist sweep_512 2>&1 | tail -2
make -B -C differentiable-mel-spectrogram_b200/csrc > /dev/null 2>&1
