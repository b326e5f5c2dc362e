#!/bin/bash
# r02 trip I (1 GPU): cfg5 dense probe on the FMA pipe; byte-set scan fast path; sanity of everything else; full bench
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02i_${name}.jsonl 2> gpurun_out/r02i_${name}.err; cut -c1-250 gpurun_out/r02i_${name}.jsonl; tail -n 2 gpurun_out/r02i_${name}.err; }
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0
run cfg1s --workload cfg1s --exps 0
run cfg2 --workload cfg2 --exps 0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02i_pytest.log 2>&1; tail -3 gpurun_out/r02i_pytest.log
timeout 900 python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; tail -c 600 gpurun_out/r02i_bench_n1.json; tail -n 3 gpurun_out/r02i_bench_n1.err
