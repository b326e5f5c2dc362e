#!/bin/bash
# r02 trip L (2 GPUs): tile distribution with prefetched super-tiles vs static split; pipelined sharded steps with the
# one-CTA-per-SM expand kernel: parity + bench N=2
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02l_${name}.jsonl 2> gpurun_out/r02l_${name}.err; cut -c1-250 gpurun_out/r02l_${name}.jsonl; tail -n 2 gpurun_out/r02l_${name}.err; }
run cfg2 --workload cfg2 --exps 0,32,0
run cfg3 --workload cfg3 --exps 0,32
run cfg5 --workload cfg5 --hay-gib 2 --steps 4 --exps 0,32
timeout 1500 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s > gpurun_out/r02l_pytest.log 2>&1
tail -9 gpurun_out/r02l_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err
tail -c 600 gpurun_out/r02l_bench_n2.json; tail -n 3 gpurun_out/r02l_bench_n2.err
