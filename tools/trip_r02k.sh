#!/bin/bash
# r02 trip K (1 GPU): global two-level tile distribution (default) vs static split (32); byte-set scan with prefetch
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02k_${name}.jsonl 2> gpurun_out/r02k_${name}.err; cut -c1-250 gpurun_out/r02k_${name}.jsonl; tail -n 2 gpurun_out/r02k_${name}.err; }
run cfg2 --workload cfg2 --exps 0,32,0
run cfg3 --workload cfg3 --exps 0,32
run cfg4 --workload cfg4 --exps 0,32
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0,32
run cfg1s --workload cfg1s --exps 0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_pytest.log 2>&1; tail -3 gpurun_out/r02k_pytest.log
