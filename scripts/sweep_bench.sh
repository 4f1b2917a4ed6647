#!/bin/bash
# BASELINE configs 2 (lambda sweep) and 3 (window-length sweep): one bench line per shape -> gpurun_out/<tag>/sweep.jsonl
TAG=${1:-sweep}
mkdir -p gpurun_out/$TAG
: > gpurun_out/$TAG/sweep.jsonl
for W in sweep_128 sweep_256 sweep_512 sweep_1024 sweep_2048 sweep_4096 esc50_10ms esc50 esc50_300ms; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline --no-e2e >> gpurun_out/$TAG/sweep.jsonl 2>> gpurun_out/$TAG/sweep.err
done
python - gpurun_out/$TAG/sweep.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(d['config']['workload'], d['config']['n_fft'], round(d['value']), round(d['ms_per_step'], 4), d['kernels_ms'],
          round(d['roofline']['step_frac'], 4))
PY
