#!/bin/bash
# One `ncu --set full` capture of the prefilter kernel for a bench workload and a set of experiment
# flags (include/acb200_debug.h), written to gpurun_out/<tag>.ncu-rep; summarise it afterwards with
#   python tools/ncu_summary.py gpurun_out/<tag>.ncu-rep
# usage (through gpurun, one GPU): bash tools/ncu_prefilter.sh <workload> <experiment flags> <tag> [hay GiB]
# A number printed by the run under ncu is never a bench value.
wl=${1:-cfg2}; exp=${2:-0}; tag=${3:-r02_prefilter_${wl}_exp${exp}}; gib=${4:-4}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:prefilter_kernel -s 2 -c 1 -f \
    -o gpurun_out/${tag} python bench.py --workload $wl --experiment $exp --hay-gib $gib --steps 1 --warmup 2 \
    --no-e2e --no-cpu-baseline --no-extras > gpurun_out/${tag}.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
