#!/bin/bash
# One GPU-box trip: parity tests, then the four single-GPU bench workloads.
# usage (through gpurun): bash tools/gpu_trip.sh [tag] [skip-tests]
tag=${1:-trip}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${tag}.log 2>&1
  tail -3 gpurun_out/pytest_${tag}.log
fi
for cfg in cfg2 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $cfg --no-cpu-baseline > gpurun_out/bench_${tag}_${cfg}.json 2> gpurun_out/bench_${tag}_${cfg}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_${cfg}.json").read().strip().splitlines()[-1])
    print("${cfg}", round(d["value"],1), "scan_ms", round(d["scan_ms"],4), "order_ms", round(d["order_ms"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "cand", d["candidates"], "matches", d["matches"])
except Exception as e:
    print("${cfg} failed", e)
PY
done
