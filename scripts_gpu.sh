#!/bin/bash
# GPU pass: parity tests, smoke, bench, per-config numbers, ncu launch list + full captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench_r01_n1.json; tail -5 gpurun_out/bench.err
python tools/bench_configs.py > gpurun_out/configs_r01.log 2>&1; echo "configs exit $?"; cut -c1-400 gpurun_out/configs_r01.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_r01_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:european_kernel -s 3 -c 1 -o gpurun_out/prof_european_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python tools/bench_configs.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pathdep_kernel -c 12 -o gpurun_out/prof_pathdep_r01 python tools/bench_configs.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu pathdep exit $?"
ls -la gpurun_out
