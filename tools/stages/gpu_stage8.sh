#!/bin/bash
# 4 GPUs: pipeline timeline (where does the ring time go?), fp8 tests + bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x -k "fp8" --timeout=300 -p no:cacheprovider > gpurun_out/test_fp8.log 2>&1; echo "fp8 tests exit $?" >> gpurun_out/summary.txt; tail -15 gpurun_out/test_fp8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 tools/pipeline_timeline.py > gpurun_out/pipe_tl4.log 2>&1; echo "timeline exit $?" >> gpurun_out/summary.txt; grep "^TL" gpurun_out/pipe_tl4.log || tail -5 gpurun_out/pipe_tl4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 4 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench_n4_bal.log 2>&1; echo "bench n4 balanced exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_n4_bal.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --dtype fp8 > gpurun_out/bench_fp8.log 2>&1; echo "bench fp8 exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/bench_fp8.log
timeout 900 python -m pytest tests/test_multigpu.py -q --timeout=800 -p no:cacheprovider -k "serve or two" > gpurun_out/test_serve.log 2>&1; echo "serve test exit $?" >> gpurun_out/summary.txt; tail -15 gpurun_out/test_serve.log
cat gpurun_out/summary.txt
