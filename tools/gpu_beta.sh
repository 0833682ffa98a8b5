#!/bin/bash
mkdir -p gpurun_out
for b in 0 1.0 0 2.0 3.0 1.0; do
  STEIN_SWEEP_BETA=$b TRACE_STEPS=20 timeout 300 python tools/step_trace.py 2>/dev/null | grep -E "median:sweep|^step" | cut -c1-100 | tr '\n' ' '; echo " beta=$b"
done
