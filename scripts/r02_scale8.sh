#!/bin/bash
# what the driver runs at round end for one N: reference arm then our arm, --steps 20 --warmup 5, under torchrun
set -u
TAG=${1:-r02v}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29541 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/ref_${N}gpu.json 2> $OUT/ref_${N}gpu.err ) 2>&1 | grep real; echo "ref rc=$?"; cut -c1-300 $OUT/ref_${N}gpu.json
( time timeout 900 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err ) 2>&1 | grep real; tail -c 800 $OUT/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("headline", d["config"]["workload"], d["n_gpus"], round(d["value"]), d["ms_per_step"], d["kernels_ms"])
    print("check", d.get("collective_check"))
    print("e2e", d.get("e2e",{}).get("value"), "pcm16", d.get("e2e_pcm16",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"))
    for k,v in d.get("extra",{}).get("workloads",{}).items(): print(k, round(v["value"]), v["ms_per_step"], v["scaling"], v["per_gpu_batch"])
except Exception as e: print("parse failed", e)
PY
