#!/bin/bash
# median tests on 1 GPU after the hparams change, then the sharded check on 8.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_mgc$N.log 2>&1; echo "mgc_rc=$?"
grep -c " ok" gpurun_out/r02_mgc$N.log; grep -i "fail\|error\|timed out" gpurun_out/r02_mgc$N.log | head -5
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "median or hint or full_size" 2>&1 | tail -3
