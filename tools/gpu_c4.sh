#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posteriors.py -m gpu -x -q -k "median or hint or full_size or engine or trajectory or posterior or deterministic" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config-e-steps 2 > gpurun_out/r02_bench_c4.log 2> gpurun_out/r02_bench_c4.err; echo "bench_rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_c4.log').read().strip().splitlines()[-1])
print("value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"])
print(l["phases_ms"]); print(l["cold"]); ce=l["config_e"]; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
PY
tail -3 gpurun_out/r02_bench_c4.err
