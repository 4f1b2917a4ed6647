#!/bin/bash
# 1-GPU visit: default bench line (large_batch headline + sub-results), parity suite, ncu captures of the three workloads'
# main kernels through the torch-free C host
set -u
TAG=${1:-r02j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
R=tests/c/abi_run; C=tests/golden/cases/timing
if [ "${SKIP_TESTS:-0}" != 1 ]; then timeout 900 python -m pytest tests -m gpu -q --timeout 90 -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log; fi
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 600 $OUT/bench_default.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("headline", d["config"]["workload"], d["value"], d["ms_per_step"], d["kernels_ms"])
print("e2e", d.get("e2e",{}).get("value"), "pcm16", d.get("e2e_pcm16",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("extra",{}).get("workloads",{}).items(): print(k, v["value"], v["ms_per_step"], v["kernels_ms"])
print("roofline", d["roofline"])
PY
for spec in "audio_mnist 256" "large_batch 1024" "esc50 128"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dmel2_(fwd|bwd)_kernel' -s 6 -c 2 -f -o $OUT/prof_$1 \
     $R time $C/$1.bin $2 3 all > $OUT/ncu_$1.log 2>&1; tail -1 $OUT/ncu_$1.log
  # the reports are ~60 MB each (gpurun_out/ is capped at 64 MiB): keep their summaries, drop the reports
  python scripts/ncu_lines.py $OUT/prof_$1.ncu-rep 30 > $OUT/ncu_$1.txt 2>&1
  python scripts/ncu_metrics.py $1=$OUT/prof_$1.ncu-rep > /dev/null 2>&1; cp profiles/ncu_metrics.json $OUT/ncu_metrics.json
  rm -f $OUT/prof_$1.ncu-rep
done
ls -la $OUT
