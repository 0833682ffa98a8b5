#!/bin/bash
# Round 2, visit 3: IMQ tests, panel kernel timing at several P budgets, ncu of the panel kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "imq or another_kernel" > gpurun_out/r02_pytest_imq.log 2>&1; echo "pytest_imq_rc=$?"
tail -8 gpurun_out/r02_pytest_imq.log
for t in 256 1024 4096; do
  echo "== STEIN_PANEL_TILES=$t"
  STEIN_PANEL_VERBOSE=1 STEIN_PANEL_TILES=$t timeout 300 python tools/panel_bench.py 32768 1024 3 2>&1 | grep -v "^$" | tail -6
done
echo "== n=65536 d=1024 tiles 1024"
STEIN_PANEL_VERBOSE=1 STEIN_PANEL_TILES=1024 timeout 300 python tools/panel_bench.py 65536 1024 2 2>&1 | tail -4
STEIN_PANEL_TILES=1024 python tools/panel_bench.py 16384 1024 1 > gpurun_out/plain_panel.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'panel_gemm_kernel' -s 2 -c 4 -f -o gpurun_out/prof_panel \
    python tools/panel_bench.py 16384 1024 1 > gpurun_out/ncu_panel.log 2>&1; echo "ncu_rc=$?"
