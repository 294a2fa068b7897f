#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|exit" gpurun_out/pytest_gpu.log | head -40
python tools/plan_sweep.py > gpurun_out/r02_plan_sweep.jsonl 2> gpurun_out/r02_plan_sweep.err; echo "sweep exit $?"; cut -c1-900 gpurun_out/r02_plan_sweep.jsonl; tail -3 gpurun_out/r02_plan_sweep.err
