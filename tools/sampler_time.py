import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
B, V = 32, 128256
torch.manual_seed(0)
logits = torch.randn(B, V, device="cuda")
out = torch.zeros(B, device="cuda", dtype=torch.int32)
seen = torch.zeros(B, (V + 31) // 32, device="cuda", dtype=torch.int32)
t = torch.full((B,), 0.7, device="cuda"); p = torch.full((B,), 0.95, device="cuda"); rp = torch.full((B,), 1.15, device="cuda")
seeds = torch.arange(B, device="cuda", dtype=torch.int32); step = torch.zeros(1, device="cuda", dtype=torch.int32)
for scale in (1.0, 4.0):
    lg = logits * scale
    for _ in range(3):
        ops.sample(lg, out, seen=seen, temperature=t, top_p=p, rep_penalty=rp, seeds=seeds, step=step)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.sample(lg, out, seen=seen, temperature=t, top_p=p, rep_penalty=rp, seeds=seeds, step=step)
    e1.record(); torch.cuda.synchronize()
    print(f"sampler B={B} V={V} logit-scale {scale}: {e0.elapsed_time(e1)*1e3/20:.1f} us")
