#!/bin/bash
# torch-free variant sweep: for every library build under build/variants/ run the C host's parity cases and timings
# Usage (under gpurun): bash scripts/r02_var.sh <tag> "<case B iters>;..." [ENV=VAL ...]
set -u
TAG=$1; SPECS=$2; shift 2
EXTRA_ENV="$*"
echo "env: $EXTRA_ENV"
OUT=gpurun_out/$TAG; mkdir -p $OUT
R=tests/c/abi_run; C=tests/golden/cases
for V in build/variants/*/; do
  VN=$(basename $V)
  echo "=== variant $VN" | tee -a $OUT/var.log
  for f in am_L512 lb_L1024 esc_L2048_norm; do LD_LIBRARY_PATH=$V timeout 120 $R check $C/$f.bin | grep -v "^check" ; done 2>&1 | tr '\n' ' ' | tee -a $OUT/var.log; echo | tee -a $OUT/var.log
  IFS=';' read -ra SP <<< "$SPECS"
  for s in "${SP[@]}"; do
    set -- $s
    env LD_LIBRARY_PATH=$V $EXTRA_ENV timeout 120 $R time $C/timing/$1.bin $2 $3 2>&1 | tee -a $OUT/var.log
  done
done
