import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.models.torch_ref import TorchPiece
from bee2bee_b200.models.weights import init_random
cfg = resolve_config("tiny-llama")
t = init_random(cfg, range(cfg.n_layers), True, True, device="cuda", dtype=torch.bfloat16, seed=0)
oracle = TorchPiece(cfg, range(cfg.n_layers), True, True, {k: v.float() for k, v in t.items()})
res = {}
LENS = (5, 16, 17, 16, 32, 64)
for graphs in (True, False):
    r = GpuRunner(cfg, "", 0, 1, torch.device("cuda:0"), max_batch=8, groups=1, max_seq_len=256, max_prefill_tokens=128,
                  seed=0, use_graphs=graphs)
    for slot, L in enumerate(LENS):
        s = SeqInit(slot=slot, prompt=[(3 * i + slot) % cfg.vocab_size for i in range(L)], pages=[1 + 2 * slot, 2 + 2 * slot],
                    temperature=0.0, top_p=1.0, repetition_penalty=1.0, seed=slot)
        r.prefill([s]); torch.cuda.synchronize()
        res[(graphs, slot)] = r.piece.logits[0, :cfg.vocab_size].clone()
    r.close()
for slot, L in enumerate(LENS):
    ids = torch.tensor([[(3 * i + slot) % cfg.vocab_size for i in range(L)]], device="cuda")
    with torch.no_grad():
        ref = oracle.forward(ids, torch.arange(L, device="cuda")[None], None)[0, -1]
    a, b = res[(True, slot)], res[(False, slot)]
    sc = ref.abs().max().item()
    print(f"slot {slot} L={L}: graph-vs-oracle {float((a-ref).abs().max())/sc:.4f}  eager-vs-oracle {float((b-ref).abs().max())/sc:.4f}  graph-vs-eager {float((a-b).abs().max())/sc:.4f}")
