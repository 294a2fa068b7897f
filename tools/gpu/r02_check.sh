#!/bin/bash
# Round-2 quick pass: GPU tests, default-h delta table, small-config latency, bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python tools/delta_default_h.py > gpurun_out/r02_delta_after.json 2> gpurun_out/r02_delta_after.err; echo "delta exit $?"; cut -c1-900 gpurun_out/r02_delta_after.json
python tools/small_configs.py 300 > gpurun_out/r02_small_after.json 2> gpurun_out/r02_small_after.err; echo "small exit $?"; cat gpurun_out/r02_small_after.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench_check.json; tail -5 gpurun_out/bench_check.err
