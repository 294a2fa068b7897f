#!/bin/bash
# Quick GPU pass: parity tests, bench line, per-config numbers (no ncu).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench exit $?"; cut -c1-3000 gpurun_out/bench_check.json; tail -5 gpurun_out/bench_check.err
python tools/bench_configs.py > gpurun_out/configs_check.log 2>&1; echo "configs exit $?"; cut -c1-330 gpurun_out/configs_check.log
