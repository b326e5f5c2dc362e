#!/bin/bash
# r02 trip A: the pending A/B of the round-1 kernel variants, in process, + ncu of the default kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 600 python tools/ab_inproc.py --workload cfg2 --exps 0,1,2,3,8,16,17,24,25,9 > gpurun_out/r02a_ab_cfg2.jsonl 2> gpurun_out/r02a_ab_cfg2.err
cat gpurun_out/r02a_ab_cfg2.jsonl | cut -c1-200
timeout 600 python tools/ab_inproc.py --workload cfg3 --exps 0,1,16,17,24,25 > gpurun_out/r02a_ab_cfg3.jsonl 2> gpurun_out/r02a_ab_cfg3.err
cat gpurun_out/r02a_ab_cfg3.jsonl | cut -c1-200
timeout 300 python tools/ab_inproc.py --workload cfg2 --engine 1 --hay-gib 1 --steps 3 --exps 0,4 > gpurun_out/r02a_ab_walk.jsonl 2> gpurun_out/r02a_ab_walk.err
cat gpurun_out/r02a_ab_walk.jsonl | cut -c1-200
timeout 400 python tools/ab_inproc.py --workload cfg5 --hay-gib 1 --steps 3 --exps 0 > gpurun_out/r02a_cfg5_host.jsonl 2> gpurun_out/r02a_cfg5_host.err
timeout 400 python tools/ab_inproc.py --workload cfg5 --hay-gib 1 --steps 3 --exps 0 --device-fill > gpurun_out/r02a_cfg5_devfill.jsonl 2> gpurun_out/r02a_cfg5_devfill.err
cat gpurun_out/r02a_cfg5_host.jsonl gpurun_out/r02a_cfg5_devfill.jsonl | cut -c1-300
timeout 600 bash tools/ncu_prefilter.sh cfg2 0 r02a_prefilter_cfg2_exp0 4
tail -3 gpurun_out/*.err
