#!/bin/bash
# Full GPU suite + smoke (what the driver runs at round end).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke_rc=$?"; tail -3 gpurun_out/r02_smoke.log
