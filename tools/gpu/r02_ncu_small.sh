#!/bin/bash
# ncu --set full of the C1 / C2 sized launches (tools/small_configs.py) with the current build.
mkdir -p gpurun_out
python tools/small_configs.py 4 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:european_kernel -c 40 -f -o /tmp/prof_small_after python tools/small_configs.py 4 > gpurun_out/ncu_small_after.log 2>&1
echo "ncu exit $?"
python tools/ncu_summary.py /tmp/prof_small_after.ncu-rep > gpurun_out/r02_ncu_small_after.txt 2>&1
grep -E "^==|gpu__time_duration|grid_size|warps_active|pipe_xu.avg" gpurun_out/r02_ncu_small_after.txt | awk 'NR<=200' | paste - - - - - | awk '{print $3, $4, $7, $11, $14, $17}' | sort | uniq -c
cp /tmp/prof_small_after.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/*.ncu-rep
