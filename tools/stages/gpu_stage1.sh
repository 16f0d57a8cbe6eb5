#!/bin/bash
# first GPU contact: probe the tcgen05 GEMM, then kernel tests per group, then model tests, smoke, tiny bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; echo "gemm_probe exit $?" >> gpurun_out/summary.txt
for grp in gemm "rmsnorm or layernorm or embed or decode_advance" attention sampler; do
  name=$(echo $grp | cut -d' ' -f1)
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "$grp" --timeout=120 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  echo "tests[$name] exit $?" >> gpurun_out/summary.txt
done
timeout 600 python -m pytest tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/test_model.log 2>&1; echo "test_model exit $?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --model tiny-llama --steps 16 --warmup 4 --batch 4 > gpurun_out/bench_tiny.log 2>&1; echo "bench_tiny exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/gemm_probe.log
