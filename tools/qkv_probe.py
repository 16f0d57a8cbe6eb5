"""Where does the QKV / O-proj GEMM lose time as the batch grows?  Variants of the same [6144 x 4096] / [4096 x 4096]
GEMM in a CUDA graph over 8 distinct weight sets: epilogue kind, inline vs precomputed 1/rms, token count, split-K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops

dev = torch.device("cuda:0")
C = ops.native(); C.init_kernels(0)
NL, H, Q, KV, D = 8, 4096, 4096, 1024, 128
g = torch.Generator(device="cuda").manual_seed(0)
wqkv = [(torch.randn(Q + 2 * KV, H, device=dev, generator=g) * 0.02).bfloat16() for _ in range(NL)]
wo = [(torch.randn(H, Q, device=dev, generator=g) * 0.02).bfloat16() for _ in range(NL)]
kc = torch.zeros(40, 64, 8, D, device=dev, dtype=torch.bfloat16); vc = torch.zeros_like(kc)

def timeit(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        def body():
            for _ in range(4):
                for l in range(NL):
                    fn(l)
        body(); s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            body()
        gr.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            gr.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * 4 * NL)

SWEEP = os.environ.get('SK_SWEEP')
for B in ((1, 32) if SWEEP else (1, 16, 17, 32)):
    x = torch.randn(B, H, device=dev).bfloat16()
    a = torch.randn(B, Q, device=dev).bfloat16()
    out_o = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    out_q = torch.empty(B, Q + 2 * KV, device=dev, dtype=torch.bfloat16)
    qb = torch.empty(B, Q, device=dev, dtype=torch.bfloat16)
    pos = torch.full((B,), 50, device=dev, dtype=torch.int32)
    slots = torch.arange(B, device=dev, dtype=torch.int32) * 64 + 64 + 50
    r = ops.rstd(x, 1e-5)
    qkv_kw = dict(epi=ops.EPI_QKV_ROPE, eps=1e-5, q_out=qb, k_cache=kc, v_cache=vc, positions=pos, slots=slots, n_q_heads=32,
                  n_kv_heads=8, head_dim=D, rope_theta=500000.0, q_scale=D ** -0.5)
    res = {}
    if SWEEP:
        for sk in (1, 2, 3, 4, 6):
            res[f'qkv rope inline splitk={sk}'] = timeit(lambda l: ops.gemm(wqkv[l], x, norm_from_x=True, splitk=sk, **qkv_kw))
        for sk in (2, 3, 4, 5, 6, 8):
            res[f'o residual splitk={sk}'] = timeit(lambda l: ops.gemm(wo[l], a, out=out_o, epi=ops.EPI_RESIDUAL, residual=x, splitk=sk))
        print(f'--- B={B} (bn={ops.pick_bn(B)})')
        for k, v in res.items():
            print(f'  {k:32s} {v:6.1f} us', flush=True)
        continue
    res["qkv rope inline-rstd"] = timeit(lambda l: ops.gemm(wqkv[l], x, norm_from_x=True, **qkv_kw))
    res["qkv rope given-rstd"] = timeit(lambda l: ops.gemm(wqkv[l], x, rstd=r, **qkv_kw))
    res["qkv rope no-norm"] = timeit(lambda l: ops.gemm(wqkv[l], x, **qkv_kw))
    res["qkv plain no-norm"] = timeit(lambda l: ops.gemm(wqkv[l], x, out=out_q))
    res["qkv plain inline-rstd"] = timeit(lambda l: ops.gemm(wqkv[l], x, out=out_q, norm_from_x=True))
    for sk in (1, 2, 4):
        res[f"qkv plain no-norm splitk={sk}"] = timeit(lambda l: ops.gemm(wqkv[l], x, out=out_q, splitk=sk))
    res["o residual"] = timeit(lambda l: ops.gemm(wo[l], a, out=out_o, epi=ops.EPI_RESIDUAL, residual=x))
    res["o plain"] = timeit(lambda l: ops.gemm(wo[l], a, out=out_o))
    for sk in (1, 2, 4, 8):
        res[f"o plain splitk={sk}"] = timeit(lambda l: ops.gemm(wo[l], a, out=out_o, splitk=sk))
    print(f"--- B={B} (bn={ops.pick_bn(B)})")
    for k, v in res.items():
        print(f"  {k:32s} {v:6.1f} us", flush=True)
