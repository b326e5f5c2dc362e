#!/bin/bash
# r02 trip N (8 GPUs): final state -- multi-rank parity at N=2 and 8 (incl. the pipelined form), bench N=8 and N=4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02n_pytest.log 2>&1
tail -10 gpurun_out/r02n_pytest.log | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02n_bench_n8.json 2> gpurun_out/r02n_bench_n8.err
tail -c 500 gpurun_out/r02n_bench_n8.json; tail -n 3 gpurun_out/r02n_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29582 \
  bench.py --gpus 4 --steps 5 --warmup 3 --no-extras > gpurun_out/r02n_bench_n4.json 2> gpurun_out/r02n_bench_n4.err
tail -c 300 gpurun_out/r02n_bench_n4.json
