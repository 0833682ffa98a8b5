#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "score or reference_run or posterior or predictions or graph" 2>&1 | tail -3
timeout 300 python tools/small_shapes_timing.py 2>&1 | tail -5 | tee gpurun_out/r02_small_shapes.log
