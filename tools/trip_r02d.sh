#!/bin/bash
# r02 trip D (1 GPU): whole -m gpu suite (incl. the new full-size cfg3/4/5 tests), cfg5 with the blocked filter, bench
mkdir -p gpurun_out
timeout 600 python tools/ab_inproc.py --workload cfg5 --hay-gib 2 --steps 4 --exps 0,32 > gpurun_out/r02d_cfg5.jsonl 2> gpurun_out/r02d_cfg5.err
cut -c1-260 gpurun_out/r02d_cfg5.jsonl; tail -n 3 gpurun_out/r02d_cfg5.err
timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r02d_pytest.log 2>&1
tail -25 gpurun_out/r02d_pytest.log
timeout 900 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
tail -c 1500 gpurun_out/r02d_bench_n1.json; tail -n 5 gpurun_out/r02d_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err
cut -c1-400 gpurun_out/r02d_bench_ref.json
