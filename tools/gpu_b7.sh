#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel" 2>&1 | tail -4
run() { echo "== $*"; env "$@" STEIN_PANEL_VERBOSE=1 STEIN_SKIP_MEDIAN=1 timeout 600 python tools/panel_bench.py 65536 1024 3 32768 2>&1 | tail -3; }
run STEIN_X=0
run STEIN_PANEL_RP=37 STEIN_PANEL_CC=52
