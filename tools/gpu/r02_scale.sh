#!/bin/bash
# Multi-GPU pass at N = $1: fused all-reduce check + latency of C1/C3/C4 (fused vs NCCL), then the bench line exactly as the driver launches it.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/fused_check.py > gpurun_out/r02_fused_check_n$N.json 2> gpurun_out/fused_check_n$N.err; echo "fused_check exit $?"; cut -c1-1500 gpurun_out/r02_fused_check_n$N.json; tail -3 gpurun_out/fused_check_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; cut -c1-400 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
