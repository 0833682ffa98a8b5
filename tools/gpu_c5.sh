#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1 0; do
STEIN_MEDIAN_PILOTLESS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e-steps 0 > gpurun_out/tmp_bench.log 2>/dev/null
python - <<PY
import json
l=json.loads(open('gpurun_out/tmp_bench.log').read().strip().splitlines()[-1])
p=l["phases_ms"]
print("pilotless=$v value %.2f ms %.3f | median %.3f sweep %.3f other %.3f | phi %.3f prep %.3f idle %.3f" % (l["value"], l["ms_per_step"], p["median"], p["sweep"], p["median"]-p["sweep"], p["phi"], p["phi_prep"], p["idle"]))
PY
done
