#!/bin/bash
# Round-2 full pass on one B200: tests, smoke, ncu launch list + full capture of the headline kernel (its summary is what
# bench.py reads roofline.traffic from), bench line, reference arm, ncu of the small / path-dependent configs.
mkdir -p gpurun_out profiles
python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --traffic-capture none"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
$B > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:european_kernel -s 3 -c 1 -f -o gpurun_out/prof_european_r02 $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py gpurun_out/prof_european_r02.ncu-rep > gpurun_out/r02_ncu_european.txt 2>&1
cp gpurun_out/r02_ncu_european.txt profiles/r02_ncu_european.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/bench.err; echo "ref exit $?"; cut -c1-900 gpurun_out/r02_bench_reference_arm.json
python tools/bench_configs.py > gpurun_out/plain3.log 2>&1; echo "configs exit $?"; cut -c1-300 gpurun_out/plain3.log
cp gpurun_out/configs_r02.json gpurun_out/r02_configs.json
ncu --set full --clock-control none -k "regex:pathdep_kernel|qmc_european_kernel|heston_kernel|jump_kernel|structured_kernel|from_normals" -c 40 -f -o /tmp/prof_other_r02 python tools/bench_configs.py --profile > gpurun_out/ncu_full2.log 2>&1
echo "ncu other exit $?"
python tools/ncu_summary.py /tmp/prof_other_r02.ncu-rep > gpurun_out/r02_ncu_other_kernels.txt 2>&1
python tools/small_configs.py 4 > /dev/null 2>&1 &&
ncu --set full --clock-control none -k regex:european_kernel -c 40 -f -o /tmp/prof_small_r02 python tools/small_configs.py 4 > gpurun_out/ncu_small.log 2>&1
python tools/ncu_summary.py /tmp/prof_small_r02.ncu-rep > gpurun_out/r02_ncu_small_configs_after.txt 2>&1
python tools/small_configs.py 300 > gpurun_out/r02_small_after.json 2>/dev/null; cat gpurun_out/r02_small_after.json
ls -la gpurun_out | tail -20; du -sh gpurun_out
