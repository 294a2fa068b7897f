#!/bin/bash
# Full GPU pass (tools/gpu/full_pass.sh): parity tests, smoke, bench, per-config numbers, ncu launch list + full captures.
# ncu reports are summarised ON the box (tools/ncu_summary.py) and only the European one travels back:
# gpurun refuses to copy more than 64 MiB of gpurun_out/.
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
ncu --set full --clock-control none --import-source on -k regex:european_kernel -s 3 -c 1 -f -o gpurun_out/prof_european_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/prof_european_r01.ncu-rep > gpurun_out/ncu_european_summary.txt 2>&1
python tools/bench_configs.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:pathdep_kernel -c 8 -f -o /tmp/prof_pathdep_r01 python tools/bench_configs.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu pathdep exit $?"
python tools/ncu_summary.py /tmp/prof_pathdep_r01.ncu-rep > gpurun_out/ncu_pathdep_summary.txt 2>&1
ncu --set full --clock-control none -k "regex:qmc_european_kernel|heston_kernel|jump_kernel|structured_kernel" -c 36 -f -o /tmp/prof_models_r01 python tools/bench_configs.py > gpurun_out/ncu_full3.log 2>&1
echo "ncu models exit $?"
python tools/ncu_summary.py /tmp/prof_models_r01.ncu-rep > gpurun_out/ncu_models_summary.txt 2>&1
ncu --set full --clock-control none -k "regex:european_from_normals_tma_kernel|from_normals_kernel" -c 8 -f -o /tmp/prof_f64_r01 python tools/bench_configs.py > gpurun_out/ncu_full4.log 2>&1
echo "ncu f64 exit $?"
python tools/ncu_summary.py /tmp/prof_f64_r01.ncu-rep > gpurun_out/ncu_f64_summary.txt 2>&1
ls -la gpurun_out; du -sh gpurun_out
