#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 64 --warmup 8 > gpurun_out/bench22_n8.log 2>&1; tail -1 gpurun_out/bench22_n8.log | cut -c1-900
