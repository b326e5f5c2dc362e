#!/bin/bash
# r02 trip F (1 GPU): batched tile draws + anchor second stage (16) A/B, byte-set scan (cfg1), K1 per-chain re-walk
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02f_${name}.jsonl 2> gpurun_out/r02f_${name}.err; cut -c1-250 gpurun_out/r02f_${name}.jsonl; tail -n 2 gpurun_out/r02f_${name}.err; }
run cfg2 --workload cfg2 --exps 0,16,32,8
run cfg3 --workload cfg3 --exps 0,16
run cfg1 --workload cfg1 --exps 0,64
run cfg4 --workload cfg4 --exps 0
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0
run walk2 --workload cfg2 --engine 1 --hay-gib 1 --steps 3 --exps 0
timeout 900 python -m pytest tests/test_gpu_zzz_experiments.py tests/test_gpu_parity.py -x -q -k "experimental or golden_find_iter or random or full_size_properties_config2" > gpurun_out/r02f_pytest.log 2>&1; tail -3 gpurun_out/r02f_pytest.log
