#!/bin/bash
# r02 trip O (2 GPUs): copy-engine payload of the begin / wait form -- multi-rank parity, then the bench at N=2
# (the warm-up calibration prints blocking vs stream ms per step), then the stream form with kernel stores for comparison
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02o_pytest.log 2>&1
tail -6 gpurun_out/r02o_pytest.log | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err
tail -c 400 gpurun_out/r02o_bench_n2.json; tail -n 3 gpurun_out/r02o_bench_n2.err
ACB_GATHER_STORE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r02o_bench_n2_store.json 2> gpurun_out/r02o_bench_n2_store.err
tail -c 300 gpurun_out/r02o_bench_n2_store.json; tail -n 3 gpurun_out/r02o_bench_n2_store.err
