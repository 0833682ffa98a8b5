#!/bin/bash
# Round 2, visit 2: panel kernels (phi and median sweep beyond 256 coordinates).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel or beyond_256" > gpurun_out/r02_pytest_panel.log 2>&1; echo "pytest_panel_rc=$?"
tail -30 gpurun_out/r02_pytest_panel.log
