#!/bin/bash
# Round 2: 8-GPU visit -- bench with timeline and the config-E sub-record.
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --config-e-steps 3 > gpurun_out/r02_bench_n$N.log 2> gpurun_out/r02_bench_n$N.err; echo "bench_n${N}_rc=$?"
cat gpurun_out/r02_bench_n$N.log; tail -5 gpurun_out/r02_bench_n$N.err
