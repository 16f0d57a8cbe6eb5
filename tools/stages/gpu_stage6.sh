#!/bin/bash
# 2 GPUs: multi-rank pipeline correctness (fused NVLink handoff) + N=2 bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -x -q --timeout=900 -p no:cacheprovider > gpurun_out/test_multigpu.log 2>&1; echo "multigpu tests exit $?" >> gpurun_out/summary.txt
tail -25 gpurun_out/test_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_n2.log
cat gpurun_out/summary.txt
