"""Does an L2-resident weight matrix make the decode GEMM faster, and does the TMA L2 prefetch get it there?
Graph of [attention -> O-proj GEMM] pairs (Llama-3-8B shapes, B=32):
  same    : every pair uses the SAME Wo (33 MB, stays in the 126 MB L2)
  cycle   : 8 different Wo (268 MB) cycled -> HBM every time, pf = 0
  cycle+pf: same, every CTA L2-prefetches its whole K slice before it waits for the attention kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.models.native import NativePiece, BatchMeta
from bee2bee_b200.models.weights import init_random

cfg = resolve_config("llama-3-8b")
NL, B = 8, 32
dev = torch.device("cuda:0")
C = ops.native(); C.init_kernels(0)
t = init_random(cfg, range(NL), False, False, device=dev, dtype=torch.bfloat16)
piece = NativePiece(cfg, range(NL), False, False, t, dev, max_tokens=64, max_seqs=64, num_pages=B + 2)
i32 = torch.int32
m = BatchMeta(ids=torch.zeros(B, device=dev, dtype=i32), positions=torch.full((B,), 20, device=dev, dtype=i32),
              slots=torch.arange(B, device=dev, dtype=i32) * 64 + 64 + 20, q_start=torch.arange(B, device=dev, dtype=i32),
              q_len=torch.ones(B, device=dev, dtype=i32), kv_len=torch.full((B,), 21, device=dev, dtype=i32),
              block_table=(torch.arange(B, device=dev, dtype=i32) + 1)[:, None].contiguous(), n_tokens=B, n_seqs=B, max_q=1)
c = cfg
x = torch.randn(64, c.hidden_size, device=dev).bfloat16()[:B]
a, x2 = piece.attn_buf[:B], piece.xb[:B]

def pair(l, pf, policy):
    ops.attention(piece.q_buf, piece.k_cache[l], piece.v_cache[l], piece.attn_buf, m.block_table, m.q_start, m.q_len,
                  m.kv_len, max_q=1, n_q=c.n_heads, n_kv=c.n_kv_heads, head_dim=c.head_dim, window=0, softcap=0.0,
                  splits=1, ws=piece.attn_ws)
    ops.gemm(piece.w[f"l{l}.wo"], a, out=x2, epi=ops.EPI_RESIDUAL, residual=x, pf=pf + policy * 4096)

for name, same, pf, policy in (("same Wo (L2 resident), evict_first loads", True, 0, 0), ("same Wo, normal-policy loads", True, 0, 1),
                               ("cycle, pf=0", False, 0, 0), ("cycle, pf=64 (TMA, up front)", False, 1024 + 64, 0),
                               ("cycle, pf=64 (LSU, up front)", False, 2048 + 64, 0), ("cycle, pf=64 TMA, normal-policy loads", False, 1024 + 64, 1),
                               ("attention only", None, 0, 0)):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        def body():
            for r in range(4):
                for l in range(NL):
                    if same is None:
                        ops.attention(piece.q_buf, piece.k_cache[l], piece.v_cache[l], piece.attn_buf, m.block_table, m.q_start,
                                      m.q_len, m.kv_len, max_q=1, n_q=c.n_heads, n_kv=c.n_kv_heads, head_dim=c.head_dim,
                                      window=0, softcap=0.0, splits=1, ws=piece.attn_ws)
                    else:
                        pair(0 if same else l, pf, policy)
        body(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s); s.synchronize()
    print(f"{name:45s}: {e0.elapsed_time(e1) * 1e3 / (5 * 4 * NL):6.2f} us per [attention + O GEMM] pair", flush=True)
