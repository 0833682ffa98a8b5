#!/bin/bash
# quick GPU check: median/phi parity tests + bench without the CPU baseline
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "median or full_size or engine or histogram" > gpurun_out/pytest_quick.log 2>&1; echo "pytest_rc=$?"
tail -5 gpurun_out/pytest_quick.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench_rc=$?"
cat gpurun_out/bench_quick.log; tail -3 gpurun_out/bench_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_quick.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_quick.log 2>&1; echo "ncu_rc=$?"
