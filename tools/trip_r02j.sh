#!/bin/bash
# r02 trip J (2 GPUs): pipelined sharded steps (begin / wait) on hardware: parity, then bench at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02j_pytest.log 2>&1
tail -12 gpurun_out/r02j_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err
tail -c 800 gpurun_out/r02j_bench_n2.json; tail -n 5 gpurun_out/r02j_bench_n2.err
