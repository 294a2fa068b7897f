#!/bin/bash
# Source-level capture of the 14-scenario European launch (C2): where its issue slots and stalls go.
mkdir -p gpurun_out
python tools/c2_once.py > gpurun_out/c2_plain.log 2>&1 || { tail -5 gpurun_out/c2_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:european_kernel -s 2 -c 1 -f -o gpurun_out/prof_c2_r02 python tools/c2_once.py > gpurun_out/ncu_c2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_c2.log; ls -la gpurun_out/prof_c2_r02.ncu-rep
