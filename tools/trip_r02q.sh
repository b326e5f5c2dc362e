#!/bin/bash
# r02 trip Q (8 GPUs): bench N=8 with the flag-closed begin / wait stream available to the warm-up calibration
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29594 \
  bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02q_bench_n8.json 2> gpurun_out/r02q_bench_n8.err
tail -c 800 gpurun_out/r02q_bench_n8.json; tail -n 3 gpurun_out/r02q_bench_n8.err
