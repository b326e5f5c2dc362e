#!/bin/bash
# r02 trip P (2 GPUs): begin / wait form closed by completion flags (no NCCL kernel in flight beside a scan)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02p_pytest.log 2>&1
tail -6 gpurun_out/r02p_pytest.log | cut -c1-260
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29593 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err
tail -c 600 gpurun_out/r02p_bench_n2.json; tail -n 3 gpurun_out/r02p_bench_n2.err
