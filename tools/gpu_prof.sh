#!/bin/bash
# Round 2 profiles: one rank's config-E phi work at several P budgets; ncu launch list of the bench; ncu --set full
# of the top kernels of config D and of the panel kernels / wide sweep.
mkdir -p gpurun_out
if [ -z "$SKIP_TUNE" ]; then
for t in 1024 1536 2048 3072; do
  echo "== one rank of config E (32768 rows x 262144 columns, d = 1024), STEIN_PANEL_TILES=$t"
  STEIN_PANEL_VERBOSE=1 STEIN_PANEL_TILES=$t STEIN_SKIP_MEDIAN=1 timeout 600 python tools/panel_bench.py 262144 1024 2 32768 2>&1 | tail -3
done
fi
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/plain_list.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo "ncu_list_rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:'flash_phi2_kernel|sweep2_tc_kernel|pair_chain_kernel|band_filter_kernel|clip_adam_kernel|err_budget_kernel|prep_x_route_kernel|prep_yt_route_kernel' -c 10 \
    -f -o gpurun_out/r02_prof_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config-e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full_rc=$?"
STEIN_SKIP_MEDIAN=0 python tools/panel_bench.py 32768 1024 1 > gpurun_out/plain_panel.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'panel_gemm_kernel' -s 1 -c 5 -f -o gpurun_out/r02_prof_panel \
    python tools/panel_bench.py 32768 1024 1 > gpurun_out/ncu_panel.log 2>&1; echo "ncu_panel_rc=$?"
