#!/bin/bash
# 8 GPUs: 4-way pipeline test, handoff micro-bench, scaling bench N=4 and N=8 (N=1,2 measured earlier)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -q --timeout=600 -p no:cacheprovider -k four > gpurun_out/test_multigpu4.log 2>&1; echo "4gpu test exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/test_multigpu4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/handoff_bench.py > gpurun_out/handoff.log 2>&1; echo "handoff exit $?" >> gpurun_out/summary.txt; grep HANDOFF gpurun_out/handoff.log || tail -5 gpurun_out/handoff.log
for N in 8 4; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 64 --warmup 8 > gpurun_out/bench_n$N.log 2>&1; echo "bench n$N exit $?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/bench_n$N.log
done
cat gpurun_out/summary.txt
