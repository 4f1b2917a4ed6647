#!/bin/bash
# One GPU-box visit: parity tests, benches, ncu launch list + full captures of the two main kernels.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [workloads...]
set -u
TAG=${1:-r01}; shift || true
WORKLOADS=${@:-audio_mnist}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for W in $WORKLOADS; do
  timeout 600 python bench.py --workload $W > $OUT/bench_$W.json 2> $OUT/bench_$W.err; tail -c 2500 $OUT/bench_$W.json; tail -3 $OUT/bench_$W.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$W.csv \
      python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_launch_$W.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dmel_(fwd|bwd)_kernel' -s 8 -c 2 -f -o $OUT/prof_$W \
      python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_full_$W.log 2>&1
  ls -la $OUT/prof_$W.ncu-rep 2>/dev/null
done
