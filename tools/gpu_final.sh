#!/bin/bash
# Round-2 record: bench on N GPUs (both arms at N = 1), the line goes to gpurun_out/r02_bench_n$N.json
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?"
  timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref_rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench_rc=$?"
fi
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print("N=$N value %.2f ms %.3f e2e %.2f" % (l["value"], l["ms_per_step"], l["e2e"]["value"]))
print(l["phases_ms"]); ce=l.get("config_e") or {}; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
PY
