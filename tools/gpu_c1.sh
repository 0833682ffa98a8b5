#!/bin/bash
# Round 2: median error-budget validation (all median tests), timings, then the full GPU suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "median or hint or full_size or engine" > gpurun_out/r02_pytest_median.log 2>&1; echo "pytest_median_rc=$?"
tail -6 gpurun_out/r02_pytest_median.log
STEIN_PANEL_VERBOSE=1 timeout 300 python tools/panel_bench.py 65536 1024 2 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e-steps 2 > gpurun_out/r02_bench_c1.log 2> gpurun_out/r02_bench_c1.err; echo "bench_rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_c1.log').read().strip().splitlines()[-1])
print("value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"])
print(l["phases_ms"]); print(l["config_e"].get("ms_per_step"), l["config_e"].get("phases_ms"), l["config_e"].get("error"))
PY
