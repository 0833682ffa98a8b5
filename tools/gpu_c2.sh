#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config-e-steps 2 > gpurun_out/r02_bench_c2.log 2> gpurun_out/r02_bench_c2.err; echo "bench_rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_c2.log').read().strip().splitlines()[-1])
print("value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"], l["e2e"].get("host_placement"))
print(l["phases_ms"]); print(l["cold"]); ce=l["config_e"]; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
PY
tail -3 gpurun_out/r02_bench_c2.err
