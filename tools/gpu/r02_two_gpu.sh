#!/bin/bash
# 2-GPU pass: sharded tests (local_devices on two GPUs, fused exchange), torchrun check of the fused all-reduce, bench line at N=2.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -rs > gpurun_out/pytest_sharded_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded_n$N.log; tail -6 gpurun_out/pytest_sharded_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/fused_check.py > gpurun_out/r02_fused_check_n$N.json 2> gpurun_out/fused_check_n$N.err; echo "fused_check exit $?"; cut -c1-1800 gpurun_out/r02_fused_check_n$N.json; tail -5 gpurun_out/fused_check_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; cut -c1-700 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
