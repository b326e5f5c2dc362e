#!/bin/bash
# r02 trip M (2 GPUs): the three tile distributions side by side (0 per-CTA default, 16 global, 32 static); bench N=2 (blocking steps)
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python tools/ab_inproc.py "$@" > gpurun_out/r02m_${name}.jsonl 2> gpurun_out/r02m_${name}.err; cut -c1-250 gpurun_out/r02m_${name}.jsonl; tail -n 2 gpurun_out/r02m_${name}.err; }
run cfg2 --workload cfg2 --exps 0,16,32,0
run cfg3 --workload cfg3 --exps 0,16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err
tail -c 600 gpurun_out/r02m_bench_n2.json; tail -n 3 gpurun_out/r02m_bench_n2.err
timeout 900 python -m pytest tests/test_gpu_zz_multirank.py -x -q -s -k "cfg2 and 2" > gpurun_out/r02m_pytest.log 2>&1; tail -4 gpurun_out/r02m_pytest.log
