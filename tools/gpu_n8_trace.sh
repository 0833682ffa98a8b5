#!/bin/bash
# N-GPU visit: step trace + bench (N = number of visible GPUs)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/step_trace.py > gpurun_out/r02_trace_n$N.log 2> gpurun_out/r02_trace_n$N.err
grep -v "^{" gpurun_out/r02_trace_n$N.log | cut -c1-170
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --config-e-steps ${2:-0} > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
p=l["phases_ms"]
print("N=$N value %.2f ms %.3f e2e %.2f prefetched %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["e2e"].get("prefetched_medians")))
print(p)
PY
