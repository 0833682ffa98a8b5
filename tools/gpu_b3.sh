#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graph_recognition or takes_a_tf1" > gpurun_out/r02_pytest_graph.log 2>&1; echo "pytest_graph_rc=$?"
tail -15 gpurun_out/r02_pytest_graph.log
