#!/bin/bash
# GPU parity suite (per-test timeout: a hung kernel costs a minute, not the visit) + smoke
set -u
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 90 -x -rA > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log; grep -c "fp64 fallback" $OUT/pytest_gpu.log
