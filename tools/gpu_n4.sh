#!/bin/bash
# 4-GPU visit: oracle check, bench with the peer-memory all-reduces, bench with NCCL all-reduces
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/mgc4.log 2>&1; echo "mgc_rc=$?"
grep -c " ok" gpurun_out/mgc4.log; grep -i "fail\|error\|timed out" gpurun_out/mgc4.log | head -5
timeout 300 $TR --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err; echo "bench_rc=$?"
tail -1 gpurun_out/bench_n4.log | cut -c1-260
STEIN_PEER_REDUCE=0 timeout 300 $TR --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4_nccl.log 2> gpurun_out/bench_n4_nccl.err; echo "bench_nccl_rc=$?"
tail -1 gpurun_out/bench_n4_nccl.log | cut -c1-260
