#!/bin/bash
# r02 trip R (1 GPU): final state -- whole -m gpu suite, bench N=1 (with configs, pinned + pageable e2e), reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02r_pytest_gpu.log | cut -c1-300
timeout 420 python bench.py > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err
tail -c 600 gpurun_out/r02r_bench_n1.json; tail -n 3 gpurun_out/r02r_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r02r_bench_reference.json 2> gpurun_out/r02r_bench_reference.err
tail -c 400 gpurun_out/r02r_bench_reference.json
