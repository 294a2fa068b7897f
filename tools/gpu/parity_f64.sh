mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_f64.py tests/test_gpu_models.py tests/test_gpu_qmc.py -q -x 2>&1 | tail -5
timeout 600 python tools/bench_configs.py 2>&1 | grep "FP64" | cut -c1-400
