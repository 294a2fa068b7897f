#!/bin/bash
# A/B of the 4-16 scenario European launches over library builds in scratch/ab/ (Philox calls in flight per thread, launch bounds)
mkdir -p gpurun_out
python tools/wide_ab.py > gpurun_out/wide_ab_base.json 2>gpurun_out/wide_ab.err; cat gpurun_out/wide_ab_base.json
for lib in scratch/ab/lib_*.so; do
  B200MC_LIB=$PWD/$lib python tools/wide_ab.py > gpurun_out/wide_ab_$(basename $lib .so).json 2>>gpurun_out/wide_ab.err; cat gpurun_out/wide_ab_$(basename $lib .so).json
done
tail -3 gpurun_out/wide_ab.err
