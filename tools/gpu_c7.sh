#!/bin/bash
# round-2 visit: prefetch test, N=1 bench, N=1 step trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "median or pilotless or hint or interleaved or prefetch or engine or trajectory or deterministic or sampler or full_size" 2>&1 | tail -5 > gpurun_out/r02_pytest_prefetch.log
cat gpurun_out/r02_pytest_prefetch.log
timeout 300 python tools/step_trace.py > gpurun_out/r02_trace_n1.log 2>&1; tail -30 gpurun_out/r02_trace_n1.log | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e-steps 0 > gpurun_out/tmp_bench.log 2>gpurun_out/tmp_bench.err
python - <<PY
import json
l=json.loads(open('gpurun_out/tmp_bench.log').read().strip().splitlines()[-1])
p=l["phases_ms"]
print("value %.2f ms %.3f | median %.3f sweep %.3f | phi %.3f prep %.3f idle %.3f e2e %.2f" % (l["value"], l["ms_per_step"], p["median"], p["sweep"], p["phi"], p["phi_prep"], p["idle"], l["e2e"]["value"]))
PY
