#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phi or engine_steps or prefetch" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print("N=1 value %.2f ms %.3f e2e %.2f frac %.3f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"]))
print(l["phases_ms"]); ce=l.get("config_e") or {}; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
PY
