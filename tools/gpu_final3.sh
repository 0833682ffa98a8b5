#!/bin/bash
# last single-GPU visit of round 2: whole GPU suite, smoke, bench (own arm), ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print("N=1 value %.2f ms %.3f e2e %.2f frac %.3f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"]))
print(l["phases_ms"]); ce=l.get("config_e") or {}; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu_list_rc=$?"
