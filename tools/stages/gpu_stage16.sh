#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 64 --warmup 8 > gpurun_out/bench16_n$n.log 2>&1; tail -1 gpurun_out/bench16_n$n.log | cut -c1-700
done
