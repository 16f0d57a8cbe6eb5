#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/tests26.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests26.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench26_n2.log 2>&1; tail -1 gpurun_out/bench26_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','p50_ttft_ms')}, d['e2e']['value'])"
cat gpurun_out/summary.txt
