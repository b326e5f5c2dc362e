#!/bin/bash
# A/B of the prefilter-kernel variants that were written without a GPU (include/acb200_debug.h:
# ACG_EXP_TALL = 1, ACG_EXP_PAIR = 2, ACG_EXP_KEY27 = 8, ACG_EXP_LOCAL2 = 16; ACG_EXP_WALK_HOT = 4 further down) against the measured default, on the bench workloads whose
# plan they apply to (stride-2 first stage with the 128 KiB bitmap: cfg2, cfg3).
# usage (through gpurun): bash tools/ab_experiments.sh [tag]
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzz_experiments.py -x -q > gpurun_out/pytest_${tag}_experiments.log 2>&1
tail -2 gpurun_out/pytest_${tag}_experiments.log
for cfg in cfg2 cfg3; do
  for exp in 0 1 2 3 8 16 17 24 25 11; do
    out=gpurun_out/bench_${tag}_${cfg}_exp${exp}.json
    timeout 600 python bench.py --workload $cfg --experiment $exp --no-cpu-baseline --no-e2e --steps 10 > $out 2> ${out%.json}.err
    python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    print("$cfg exp $exp scan_ms", round(d["scan_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "cand", d["candidates"], "matches", d["matches"])
except Exception as e:
    print("$cfg exp $exp failed", e)
PY
  done
done
# walk engine (K1): plain vs hot rows in shared memory
for exp in 0 4; do
  out=gpurun_out/bench_${tag}_cfg2_walk_exp${exp}.json
  timeout 600 python bench.py --workload cfg2 --engine 1 --experiment $exp --no-cpu-baseline --no-e2e --hay-gib 1 --steps 5 > $out 2> ${out%.json}.err
  python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    print("cfg2 walk exp $exp scan_ms", round(d["scan_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "matches", d["matches"])
except Exception as e:
    print("cfg2 walk exp $exp failed", e)
PY
done
# dense table built on the host (+ H2D) vs on the device: build_s of the 100 000-pattern automaton
for flag in "" "--device-fill"; do
  out=gpurun_out/bench_${tag}_cfg5_build${flag:+_devfill}.json
  timeout 900 python bench.py --workload cfg5 $flag --no-cpu-baseline --no-e2e --hay-gib 1 --steps 3 > $out 2> ${out%.json}.err
  python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    print("cfg5 build ${flag:-host}", "build_s", round(d["build_s"], 3), "scan_ms", round(d["scan_ms"], 4), "matches", d["matches"])
except Exception as e:
    print("cfg5 build ${flag:-host} failed", e)
PY
done
