#!/bin/bash
# N-GPU visit (N = $1, default 2): oracle check of the sharded engine, step trace, bench
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MGC_QUICK=${MGC_QUICK:-1} timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_mgc$N.log 2>&1; echo "mgc_rc=$?"
grep -c " ok" gpurun_out/r02_mgc$N.log; grep -i "fail\|error\|timed out" gpurun_out/r02_mgc$N.log | head -5; grep "prefetched" gpurun_out/r02_mgc$N.log
timeout 300 $TR --master-port 29512 tools/step_trace.py > gpurun_out/r02_trace_n$N.log 2> gpurun_out/r02_trace_n$N.err
grep -v "^{" gpurun_out/r02_trace_n$N.log | cut -c1-150
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --config-e-steps ${2:-0} > gpurun_out/tmp_bench_n$N.log 2> gpurun_out/tmp_bench_n$N.err; echo "bench_rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/tmp_bench_n$N.log').read().strip().splitlines()[-1])
p=l["phases_ms"]
print("N=$N value %.2f ms %.3f e2e %.2f prefetched %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("prefetched_medians")))
print(p)
PY
