#!/bin/bash
# N=1 evidence refresh: smoke, bench (both arms), ncu launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench_rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref_rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu_list_rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'flash_phi2_kernel|sweep2_tc_kernel|pair_chain_kernel|pilot_h16_kernel|band_filter_kernel|clip_adam_kernel' -c 7 \
    -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu_full_rc=$?"
tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-300
