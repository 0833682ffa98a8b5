#!/bin/bash
STEIN_MEDIAN_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "pilotless" 2>&1 | grep -E "direct_ok|passed|failed|Assert" | head -20
