#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_multigpu.py -q --timeout=200 -k "half_layer" -p no:cacheprovider > gpurun_out/tests18a.log 2>&1; echo "half-layer tests exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/tests18a.log
timeout 600 python -m pytest tests/test_multigpu.py -q --timeout=300 -k "not half_layer" -p no:cacheprovider > gpurun_out/tests18b.log 2>&1; echo "multigpu tests exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests18b.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench18_n2.log 2>&1; tail -1 gpurun_out/bench18_n2.log | cut -c1-600
cat gpurun_out/summary.txt
