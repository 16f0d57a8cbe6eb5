# Multi-GPU measurement session (run on an 8-GPU box: `gpurun --gpus 8 -- bash tools/n8_session.sh [full]`).
# Default: the 8-GPU correctness test, the headline at N = 8 and N = 4, mxfp8 at N = 8.
# `full` adds the constructed NCCL arm, configs 2 / 4 / 5, the per-rank stage timeline and the HTTP load test.
set -u
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
OUT=gpurun_out/n8
mkdir -p $OUT
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  $R --nproc-per-node $np --master-port $((29800 + RANDOM % 100)) bench.py --gpus $np "$@" > $OUT/$name.log 2>&1
  echo "== $name rc=$?"; grep "^{" $OUT/$name.log | tail -1 | cut -c1-2200
}
python -m pytest tests/test_multigpu.py -x -q -m gpu -k "eight or four" 2>&1 | tail -3
run ours_n8 8 --steps 20 --warmup 5
run ours_n4 4 --steps 20 --warmup 5
run cfg3_ours_n8 8 --config 3 --steps 20 --warmup 5
# rank-count independence: the greedy checksum of the bench line must be the same at every N
python bench.py --steps 8 --warmup 3 > $OUT/ours_n1.log 2>&1
grep -ho '"greedy_check": {[^}]*}' $OUT/ours_n1.log $OUT/ours_n4.log $OUT/ours_n8.log
if [ "${1:-}" = "full" ]; then
  B2B_MX_HANDOFF=0 run cfg3_ours_n8_bf16hop 8 --config 3 --steps 20 --warmup 5      # A/B of the quantised hop
  run nccl_n8 8 --steps 20 --warmup 5 --impl nccl
  run cfg2_ours_n8 8 --config 2 --steps 64 --warmup 8
  run cfg2_nccl_n8 8 --config 2 --steps 64 --warmup 8 --impl nccl --no-e2e
  run cfg5_ours_n8 8 --config 5 --steps 20 --warmup 5
  run cfg4_ours_n4 4 --config 4
  run cfg4_nccl_n4 4 --config 4 --impl nccl --no-e2e
  $R --nproc-per-node 8 --master-port 29791 tools/pipeline_timeline.py 2>&1 | grep "^TL" | tee $OUT/timeline.log
  python tools/load_test.py --spawn zephyr-7b-beta --pieces 8 --max-batch 64 --clients 32 --requests 2 --max-new-tokens 48 2>&1 | tail -5 | tee $OUT/load_test.log
fi
