#!/bin/bash
# tile-shape sweep with the torch-free C host: DMEL_NG x DMEL_SF per workload
set -u
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
R=tests/c/abi_run; C=tests/golden/cases/timing
sw() { # case B iters "ng list" "sf list"
  for ng in $4; do for sf in $5; do
    echo -n "$1 NG=$ng SF=$sf: " | tee -a $OUT/sweep.log
    DMEL_NG=$ng DMEL_SF=$sf timeout 60 $R time $C/$1.bin $2 $3 2>&1 | sed 's/.*mels=[0-9]*: //' | tee -a $OUT/sweep.log
  done; done
}
sw audio_mnist 256 200 "4 8" "2 4 8"
sw large_batch 1024 8 "2 4" "2 4 8"
sw esc50 128 20 "1 2" "4 8 16"
