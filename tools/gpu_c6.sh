#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine or trajectory or deterministic or hint or pilotless or full_size or sampler or smoke" 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e-steps 0 > gpurun_out/tmp_bench.log 2>/dev/null
python - <<PY
import json
l=json.loads(open('gpurun_out/tmp_bench.log').read().strip().splitlines()[-1])
p=l["phases_ms"]
print("value %.2f ms %.3f | median %.3f sweep %.3f other %.3f | phi %.3f prep %.3f idle %.3f e2e %.2f" % (l["value"], l["ms_per_step"], p["median"], p["sweep"], p["median"]-p["sweep"], p["phi"], p["phi_prep"], p["idle"], l["e2e"]["value"]))
PY
done
