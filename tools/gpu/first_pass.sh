#!/bin/bash
# First GPU pass: parity tests, smoke, bench, launch list, one full ncu capture of the top kernel.
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_first.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench_r01_first.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_first.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref_first.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:european_kernel -s 3 -c 1 -o gpurun_out/prof_european_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
