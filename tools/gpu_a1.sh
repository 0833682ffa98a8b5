#!/bin/bash
# Round 2, visit 1: conditioning study, new parity tests, smoke, a short bench.
mkdir -p gpurun_out
python tools/phi_conditioning_study.py > gpurun_out/r02_cond_study.log 2>&1; echo "study_rc=$?"
python tools/trunc_study.py > gpurun_out/r02_trunc_study.log 2>&1; echo "trunc_rc=$?"; cat gpurun_out/r02_trunc_study.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest_rc=$?"
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke_rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a1.log 2> gpurun_out/r02_bench_a1.err; echo "bench_rc=$?"
cat gpurun_out/r02_cond_study.log; tail -5 gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_smoke.log; cat gpurun_out/r02_bench_a1.log
