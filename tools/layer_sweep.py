"""Times one Llama-3-8B decoder layer (5 launches) inside a CUDA graph for split-K / PDL variants.
8 distinct layers' weights (3.5 GB >> L2) are cycled so every GEMM streams from HBM."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.models.native import NativePiece, BatchMeta
from bee2bee_b200.models.weights import init_random

cfg = resolve_config("llama-3-8b")
NL = 8
dev = torch.device("cuda:0")
C = ops.native(); C.init_kernels(0)
t = init_random(cfg, range(NL), False, False, device=dev, dtype=torch.bfloat16)
results = []
for B in [int(os.environ.get("B", "32"))]:
    piece = NativePiece(cfg, range(NL), False, False, t, dev, max_tokens=64, max_seqs=64, num_pages=B + 2)
    i32 = torch.int32
    meta = BatchMeta(ids=torch.zeros(B, device=dev, dtype=i32), positions=torch.full((B,), 20, device=dev, dtype=i32),
                     slots=torch.arange(B, device=dev, dtype=i32) * 64 + 64 + 20, q_start=torch.arange(B, device=dev, dtype=i32),
                     q_len=torch.ones(B, device=dev, dtype=i32), kv_len=torch.full((B,), 21, device=dev, dtype=i32),
                     block_table=(torch.arange(B, device=dev, dtype=i32) + 1)[:, None].contiguous(), n_tokens=B, n_seqs=B, max_q=1)
    x = torch.randn(64, cfg.hidden_size, device=dev).bfloat16()
    H, F, Q, KV = cfg.hidden_size, cfg.ffn_size, cfg.q_dim, cfg.kv_dim
    shapes = {"qkv": (Q + 2 * KV, H), "o": (H, Q), "gu": (2 * F, H), "down": (H, F)}
    combos = [dict(qkv=4, o=4, gu=1, down=8), dict(qkv=2, o=4, gu=1, down=4), dict(qkv=4, o=8, gu=2, down=8),
              dict(qkv=2, o=2, gu=1, down=4), dict(qkv=4, o=4, gu=2, down=4), dict(qkv=1, o=1, gu=1, down=1),
              dict(qkv=8, o=8, gu=1, down=8), dict(qkv=2, o=4, gu=2, down=8)]
    if os.environ.get('QUICK'):
        combos = [dict(qkv=2, o=4, gu=1, down=4), dict(qkv=4, o=4, gu=1, down=4), dict(qkv=4, o=4, gu=1, down=8), dict(qkv=4, o=8, gu=1, down=8),
                  dict(qkv=2, o=4, gu=2, down=4), dict(qkv=4, o=4, gu=2, down=8), dict(qkv=2, o=2, gu=1, down=4), dict(qkv=2, o=4, gu=1, down=8)]
    for pdl in ((True,) if os.environ.get('QUICK') else (True, False)):
        C.set_pdl(pdl)
        for cb in combos:
            ops.SPLITK_OVERRIDE.clear()
            for k, v in cb.items():
                ops.SPLITK_OVERRIDE[shapes[k]] = v
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                piece.forward(meta, x_in=x)           # warm
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(4):
                        piece.forward(meta, x_in=x)   # 32 layers
                g.replay(); s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(5):
                    g.replay()
                e1.record(s); s.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (5 * 4 * NL)
            wb = sum(a * b * 2 for a, b in shapes.values())
            print(f"B={B} pdl={int(pdl)} {cb}: {us:7.1f} us/layer  ({wb / us / 1e3:6.0f} GB/s)", flush=True)
    del piece
