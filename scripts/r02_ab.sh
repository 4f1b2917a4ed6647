#!/bin/bash
# Round-2 A/B visit: GPU parity suite on the second-generation kernels, then bench lines of both generations.
# Usage (under gpurun): bash scripts/r02_ab.sh <tag> [workloads...]
set -u
TAG=${1:-r02a}; shift || true
WORKLOADS=${@:-audio_mnist}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -15 $OUT/pytest_gpu.log
for W in $WORKLOADS; do
  for GEN in 2 1; do
    V1=0; [ $GEN = 1 ] && V1=1
    DMEL_V1=$V1 timeout 600 python bench.py --workload $W --no-cpu-baseline --no-e2e > $OUT/bench_${W}_gen$GEN.json 2> $OUT/bench_${W}_gen$GEN.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${W}_gen$GEN.json").read().strip().splitlines()[-1])
    print("$W gen$GEN", d["value"], d["ms_per_step"], d.get("kernels_ms"), d.get("lambda_only"))
except Exception as e:
    print("$W gen$GEN failed", e); print(open("$OUT/bench_${W}_gen$GEN.err").read()[-1500:])
PY
  done
done
