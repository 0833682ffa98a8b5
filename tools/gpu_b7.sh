#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel" 2>&1 | tail -4
run() { echo "== $*"; env "$@" STEIN_PANEL_VERBOSE=1 STEIN_SKIP_MEDIAN=1 timeout 600 python tools/panel_bench.py 65536 1024 2 32768 2>&1 | tail -2; }
run STEIN_PANEL_FUSED=1
run STEIN_PANEL_FUSED=1 STEIN_PANEL_RP=18 STEIN_PANEL_CC=8
run STEIN_PANEL_FUSED=1 STEIN_PANEL_RP=37 STEIN_PANEL_CC=4
run STEIN_PANEL_FUSED=1 STEIN_PANEL_RP=37 STEIN_PANEL_CC=6
run STEIN_PANEL_FUSED=1 STEIN_PANEL_RP=18 STEIN_PANEL_CC=12
run STEIN_PANEL_FUSED=0
