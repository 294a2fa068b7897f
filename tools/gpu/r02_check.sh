#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|exit|^E  " gpurun_out/pytest_gpu.log | head -30
python tools/rng_evidence.py > gpurun_out/r02_rng_evidence.json 2> gpurun_out/rng_evidence.err; echo "rng exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_rng_evidence.json')); s=d['stream_statistics_1.07e10_draws']
print({k:s[k] for k in ('draws','mean','second_moment','fourth_moment','same_word','lag1','chi2_z','chi2_z_all_256_bins_vs_normal_law','tails')}); print(d['stream_statistics_5.4e8_draws']['chi2_joint_64x64'])
for r in d['strike_sweep_single_step_2^32_samples']:
    if abs(r['k_sigma'])>=3.5 or r['k_sigma']==0: print(r['k_sigma'], r['type'], '%.4e'%r['price'], '%.4e'%r['bs'], 'z=%.2f'%r['z'], 'rel=%.4f'%r['rel'])
"
python tools/floor_probe.py > gpurun_out/r02_floor_probe.json 2>&1; cat gpurun_out/r02_floor_probe.json
python tools/bench_configs.py 2>/dev/null | grep -E "QMC" | cut -c1-260
python tools/sanitize_probe.py 2>&1 | tail -1 | cut -c1-300
