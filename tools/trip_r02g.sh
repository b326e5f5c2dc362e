#!/bin/bash
# r02 trip G (1 GPU): cfg5 with two anchor lookups in flight per lane + sparser anchor map; byte-set scan on cfg1s; K1 restored
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02g_${name}.jsonl 2> gpurun_out/r02g_${name}.err; cut -c1-250 gpurun_out/r02g_${name}.jsonl; tail -n 2 gpurun_out/r02g_${name}.err; }
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0
run cfg1s --workload cfg1s --exps 0,64
run cfg1 --workload cfg1 --exps 0
run cfg2 --workload cfg2 --exps 0
run cfg3 --workload cfg3 --exps 0
run walk2 --workload cfg2 --engine 1 --hay-gib 1 --steps 3 --exps 0
timeout 600 bash tools/ncu_prefilter.sh cfg5 0 r02g_prefilter_cfg5 2
