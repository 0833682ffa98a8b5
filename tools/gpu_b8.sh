#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel or beyond_256" 2>&1 | tail -4
timeout 600 python tools/panel_bench.py 65536 1024 2 2>&1 | tail -3
