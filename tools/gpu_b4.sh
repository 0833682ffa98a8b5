#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel" > gpurun_out/r02_pytest_panel.log 2>&1; echo "pytest_panel_rc=$?"
tail -6 gpurun_out/r02_pytest_panel.log
STEIN_PANEL_VERBOSE=1 STEIN_SKIP_MEDIAN=1 timeout 600 python tools/panel_bench.py 262144 1024 3 32768 2>&1 | tail -5
STEIN_SKIP_MEDIAN=1 python tools/panel_bench.py 32768 1024 1 > gpurun_out/plain_panel.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'panel_gemm_kernel' -s 1 -c 4 -f -o gpurun_out/r02_prof_panel \
    python tools/panel_bench.py 32768 1024 1 > gpurun_out/ncu_panel.log 2>&1; echo "ncu_panel_rc=$?"
