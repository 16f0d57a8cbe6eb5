"""Llama-3-8B batch-32 decode without CUDA graphs so that ncu sees individual launches."""
import os, sys
os.environ.setdefault("B2B_ALLOW_RANDOM_WEIGHTS", "1")     # no checkpoints offline: random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config

model = os.environ.get("B2B_MODEL", "llama-3-8b")
B = int(os.environ.get("B2B_BATCH", "32"))
steps = int(os.environ.get("B2B_STEPS", "3"))
cfg = resolve_config(model)
r = GpuRunner(cfg, "", 0, 1, torch.device("cuda:0"), max_batch=B, groups=1, max_seq_len=512, max_prefill_tokens=512,
              use_graphs=False)
seqs = [SeqInit(slot=i, prompt=[(7 + 131 * i + 31 * j) % 100000 + 256 for j in range(16)], pages=[1 + 2 * i, 2 + 2 * i],
                temperature=0.7, seed=i) for i in range(B)]
r.prefill(seqs)
r.decode(2)
r.sync()
torch.cuda.profiler.start()
r.decode(steps)
r.sync()
torch.cuda.profiler.stop()
print("done", r.kernel_launches)
