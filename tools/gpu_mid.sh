#!/bin/bash
# full GPU test suite + bench without the CPU baseline + ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench_rc=$?"
cat gpurun_out/bench_quick.log; tail -3 gpurun_out/bench_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_quick.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_quick.log 2>&1; echo "ncu_rc=$?"
