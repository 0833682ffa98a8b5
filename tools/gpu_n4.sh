#!/bin/bash
# N-GPU visit: oracle check of the sharded path, bench with timeline + config E.
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_mgc$N.log 2>&1; echo "mgc_rc=$?"
grep -c " ok" gpurun_out/r02_mgc$N.log; grep -i "fail\|error\|timed out" gpurun_out/r02_mgc$N.log | head -5
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --config-e-steps 2 > gpurun_out/r02_bench_n$N.log 2> gpurun_out/r02_bench_n$N.err; echo "bench_rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n$N.log').read().strip().splitlines()[-1])
print("value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"])
print(l["phases_ms"]); print(l["config_e"].get("ms_per_step"), l["config_e"].get("phases_ms"), l["config_e"].get("error"))
PY
