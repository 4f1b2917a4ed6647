#!/bin/bash
# N-GPU visit: the >=2-GPU parity test, the default bench line under torchrun (headline + sub-results + collective_check),
# the reference arm under torchrun, and repeated short runs that must all exit normally (teardown check).
# Usage (under gpurun --gpus N): bash scripts/r02_multi.sh <tag> <N> [repeats]
set -u
TAG=${1:-r02n}; N=${2:-2}; REP=${3:-5}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 -rA > $OUT/pytest_multi.log 2>&1; tail -4 $OUT/pytest_multi.log
timeout 900 $TR --master-port 29511 bench.py --gpus $N > $OUT/bench_default_${N}gpu.json 2> $OUT/bench_default_${N}gpu.err; echo "default rc=$?"
tail -c 1500 $OUT/bench_default_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_default_${N}gpu.json").read().strip().splitlines()[-1])
    print("headline", d["config"]["workload"], d["n_gpus"], d["value"], d["ms_per_step"], d["kernels_ms"])
    print("collective", d["config"]["collective"][:60]); print("check", d.get("collective_check"))
    print("e2e", d.get("e2e",{}).get("value"), "pcm16", d.get("e2e_pcm16",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"))
    for k,v in d.get("extra",{}).get("workloads",{}).items(): print(k, v["value"], v["ms_per_step"], v["scaling"], v["per_gpu_batch"])
except Exception as e: print("parse failed", e)
PY
timeout 300 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_reference_${N}gpu.json 2> $OUT/bench_reference_${N}gpu.err; echo "reference rc=$?"; head -c 400 $OUT/bench_reference_${N}gpu.json; echo
ok=0
for i in $(seq 1 $REP); do
  for C in side fused nccl; do
    timeout 200 $TR --master-port $((29520+i)) bench.py --gpus $N --workload audio_mnist --no-cpu-baseline --no-e2e --collective $C > $OUT/rep_${C}_$i.json 2> $OUT/rep_${C}_$i.err
    rc=$?; [ $rc = 0 ] && ok=$((ok+1))
    python - <<PY >> $OUT/teardown.log
import json
try:
    d=json.loads(open("$OUT/rep_${C}_$i.json").read().strip().splitlines()[-1])
    print("run $i $C rc=$rc", round(d["value"]), round(d["ms_per_step"],5), d.get("collective_check"))
except Exception as e: print("run $i $C rc=$rc parse failed", e)
PY
  done
done
echo "normal exits: $ok of $((3*REP))" | tee -a $OUT/teardown.log
cat $OUT/teardown.log | tail -20
