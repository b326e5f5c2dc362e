#!/bin/bash
# r02 trip E (1 GPU): K1 with four chains per lane; ncu --set full of the default prefilter kernel on cfg2 and cfg5
mkdir -p gpurun_out
timeout 300 python tools/ab_inproc.py --workload cfg2 --engine 1 --hay-gib 1 --steps 3 --exps 0 > gpurun_out/r02e_walk_cfg2.jsonl 2> gpurun_out/r02e_walk.err
timeout 300 python tools/ab_inproc.py --workload cfg5 --engine 1 --hay-gib 1 --steps 3 --exps 0 > gpurun_out/r02e_walk_cfg5.jsonl 2>> gpurun_out/r02e_walk.err
cut -c1-260 gpurun_out/r02e_walk_cfg2.jsonl gpurun_out/r02e_walk_cfg5.jsonl; tail -n 3 gpurun_out/r02e_walk.err
timeout 600 bash tools/ncu_prefilter.sh cfg2 0 r02e_prefilter_cfg2 4
timeout 600 bash tools/ncu_prefilter.sh cfg5 0 r02e_prefilter_cfg5 2
ncu --set full --clock-control none --import-source on -k regex:walk_overlapping -s 1 -c 1 -f -o gpurun_out/r02e_walk_cfg2 \
  python bench.py --workload cfg2 --engine 1 --hay-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02e_walk_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
