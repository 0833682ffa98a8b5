#!/bin/bash
# Round-2 final single-GPU record: the whole GPU suite, smoke, both bench arms, ncu launch list, ncu --set full of the
# kernels of one iteration, ncu --set full of the score kernels at the example shapes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref_rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print("N=1 value %.2f ms %.3f e2e %.2f frac %.3f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"]))
print(l["phases_ms"]); ce=l.get("config_e") or {}; print(ce.get("ms_per_step"), ce.get("phases_ms"), ce.get("roofline",{}).get("frac"), ce.get("error"))
print(l.get("cpu_baseline"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu_list_rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'flash_phi2_kernel|sweep2_tc_kernel|pair_chain_kernel|band_filter_kernel|clip_adam_kernel|x_stats_kernel|split_f16_kernel|prep_x_route_kernel|prep_yt_route_kernel|colmax_sx_partial_kernel|finalize_slots_kernel' -c 11 \
    -f -o gpurun_out/r02_prof_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full_rc=$?"
SMALL_WARM=1 SMALL_ITERS=2 ncu --set full --clock-control none --import-source on -k regex:'score_kernel' -c 12 \
    -f -o gpurun_out/r02_prof_scores python tools/small_shapes_timing.py > gpurun_out/ncu_scores.log 2>&1; echo "ncu_scores_rc=$?"
ls -la gpurun_out/*.ncu-rep
