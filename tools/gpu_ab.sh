#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer" 2>&1 | grep -E "assert|Error|passed|failed" | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "prefetch or interleaved or pilotless or engine_steps" 2>&1 | tail -4
for v in 0 1 0 1; do
  echo "== STEIN_DEVICE_BW=$v"
  STEIN_DEVICE_BW=$v TRACE_STEPS=20 timeout 300 python tools/step_trace.py 2>/dev/null | grep -v "^{" | cut -c1-75
done
