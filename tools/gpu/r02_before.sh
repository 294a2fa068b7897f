#!/bin/bash
# Round-2 "before" pass with the round-1 build: the small-bump delta defect on hardware, C1/C2 API latency, ncu of the C1/C2 launches.
mkdir -p gpurun_out
python tools/delta_default_h.py > gpurun_out/r02_delta_before.json 2> gpurun_out/r02_delta_before.err; echo "delta exit $?"; cut -c1-1500 gpurun_out/r02_delta_before.json
python tools/small_configs.py 300 > gpurun_out/r02_small_before.json 2> gpurun_out/r02_small_before.err; echo "small exit $?"; cat gpurun_out/r02_small_before.json
python tools/small_configs.py 4 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:european_kernel -c 40 -f -o /tmp/prof_small_before python tools/small_configs.py 4 > gpurun_out/ncu_small_before.log 2>&1
echo "ncu exit $?"
python tools/ncu_summary.py /tmp/prof_small_before.ncu-rep > gpurun_out/r02_ncu_small_before.txt 2>&1
grep -c "==" gpurun_out/r02_ncu_small_before.txt
