#!/bin/bash
# Multi-GPU bench line exactly as the driver launches it (N = $1).
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r01_n$N.json 2> gpurun_out/bench_n$N.err
echo "N=$N exit $?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_r01_n$N.json | cut -c1-600; tail -3 gpurun_out/bench_n$N.err
