#!/bin/bash
# r02 trip B: dynamic tile distribution (ACG_EXP_DYN = 32) x 27-bit keys (8) x lane-local second stage (16)
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02b_${name}.jsonl 2> gpurun_out/r02b_${name}.err
  cut -c1-230 gpurun_out/r02b_${name}.jsonl
}
run cfg2 --workload cfg2 --exps 0,8,32,40,48,56
run cfg3 --workload cfg3 --exps 0,8,24,32,40,56
run cfg4 --workload cfg4 --exps 0,8,32,40
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0,32
run cfg2_1g --workload cfg2 --hay-gib 1 --exps 0,32
run cfg2_2g --workload cfg2 --hay-gib 2 --exps 0,32
for f in gpurun_out/r02b_*.err; do tail -n 2 $f; done
