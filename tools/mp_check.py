"""Multi-rank pipeline check (run under torchrun): N pieces over N GPUs with the fused
NVLink handoff must produce exactly the tokens of the single-GPU engine."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.parallel.dist import init_distributed, shutdown

model = os.environ.get("B2B_MODEL", "tiny-llama")
steps = int(os.environ.get("B2B_STEPS", "12"))
groups = int(os.environ.get("B2B_GROUPS", "0"))
B = int(os.environ.get("B2B_BATCH", "4"))
rank, world, local = init_distributed()
cfg = resolve_config(model)
groups = groups or world
total = B * groups
r = GpuRunner(cfg, "", rank, world, torch.device(f"cuda:{local}"), max_batch=total, groups=groups, max_seq_len=256,
              max_prefill_tokens=128, seed=0)
V = cfg.vocab_size
seqs = [SeqInit(slot=i, prompt=[(13 * i + 7 * j + 5) % (V - 8) + 4 for j in range(5 + 3 * i)], pages=[1 + 4 * i, 2 + 4 * i, 3 + 4 * i, 4 + 4 * i],
                temperature=0.0, top_p=1.0, repetition_penalty=1.0, seed=i) for i in range(total)]
r.prefill(seqs)
r.decode(steps // 2)
r.sync()
r.decode(steps - steps // 2)          # second burst: flags are re-armed between bursts
r.sync()
if rank == 0:
    hist, _ = r.read_history()
    print("RESULT " + json.dumps({"world": world, "tokens": hist[:total, :steps + 1].tolist()}), flush=True)
r.close()
shutdown()
