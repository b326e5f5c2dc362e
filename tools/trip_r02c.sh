#!/bin/bash
# r02 trip C (2 GPUs): the sharded C-ABI path on hardware -- parity tests, then bench at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02c_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02c_pytest.log 2>&1
tail -15 gpurun_out/r02c_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -c 3000 gpurun_out/r02c_bench_n2.json
tail -5 gpurun_out/r02c_bench_n2.err
