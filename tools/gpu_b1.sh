#!/bin/bash
# Round 2, visit 2: panel kernels (phi and median sweep beyond 256 coordinates), posterior tests, timeline bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "panel or beyond_256" > gpurun_out/r02_pytest_panel.log 2>&1; echo "pytest_panel_rc=$?"
tail -25 gpurun_out/r02_pytest_panel.log
timeout 900 python -m pytest tests/test_gpu_posteriors.py -m gpu -q -s > gpurun_out/r02_pytest_post.log 2>&1; echo "pytest_post_rc=$?"
tail -12 gpurun_out/r02_pytest_post.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b1.log 2> gpurun_out/r02_bench_b1.err; echo "bench_rc=$?"
cat gpurun_out/r02_bench_b1.log; tail -5 gpurun_out/r02_bench_b1.err
