import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config
cfg = resolve_config("tiny-llama")
def run(graphs):
    r = GpuRunner(cfg, "", 0, 1, torch.device("cuda:0"), max_batch=8, groups=1, max_seq_len=256, max_prefill_tokens=128, seed=0, use_graphs=graphs)
    out = []
    for slot, L in enumerate((16, 16)):
        prompt = [(3 * i + slot) % cfg.vocab_size for i in range(L)]
        s = SeqInit(slot=slot, prompt=prompt, pages=[1 + 2 * slot, 2 + 2 * slot], temperature=0.0, top_p=1.0, repetition_penalty=1.0, seed=slot)
        r.prefill([s]); torch.cuda.synchronize()
        p = r.piece
        out.append({k: v.float().clone() for k, v in dict(xa=p.xa[:16], xb=p.xb[:16], q=p.q_buf[:16], attn=p.attn_buf[:16], h=p.h_buf[:16],
                                                          last_x=p.last_x[:1], logits=p.logits[0, :cfg.vocab_size],
                                                          k0=p.k_cache[0][1 + 2 * slot, :16], k3=p.k_cache[3][1 + 2 * slot, :16]).items()})
    r.close()
    return out
g, e = run(True), run(False)
print("env STREAMK", os.environ.get("B2B_STREAMK"), "PDL", os.environ.get("B2B_PDL"))
for i in range(2):
    print(f" call {i}: " + "  ".join(f"{k}:{float((g[i][k]-e[i][k]).abs().max()):.4f}" for k in g[i]))
