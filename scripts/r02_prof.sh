#!/bin/bash
# torch-free tuning visit: parity of the C host on the committed case files, timings, one ncu --set full capture
# Usage (under gpurun): bash scripts/r02_prof.sh <tag> [ncu_case B]
set -u
TAG=${1:-r02b}; NCASE=${2:-audio_mnist}; NB=${3:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
R=tests/c/abi_run; C=tests/golden/cases
for f in am_L512 lb_L1024 esc_L2048_norm; do timeout 120 $R check $C/$f.bin; done 2>&1 | tee $OUT/check.log
( timeout 120 $R time $C/timing/audio_mnist.bin 256 200
  timeout 120 $R time $C/timing/large_batch.bin 1024 10
  timeout 120 $R time $C/timing/esc50.bin 128 20 ) 2>&1 | tee $OUT/time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dmel2?_(fwd|bwd)_kernel' -s 6 -c 2 -f -o $OUT/prof_$NCASE \
   $R time $C/timing/$NCASE.bin $NB 3 all > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log; ls -la $OUT
