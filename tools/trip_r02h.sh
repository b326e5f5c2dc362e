#!/bin/bash
# r02 trip H (8 GPUs): multi-rank parity at N=8 (cfg2, cfg5), bench at N=8 (cfg2 + BASELINE config 5: 100 000 patterns / 32 GiB)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02h_pytest.log 2>&1
tail -12 gpurun_out/r02h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err
tail -c 1200 gpurun_out/r02h_bench_n8.json; tail -n 5 gpurun_out/r02h_bench_n8.err
