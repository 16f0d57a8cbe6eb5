"""Small-prefill GEMMs (T = 256 / 512 / 1024 tokens) on the Llama-3-8B shapes whose weight-tile count cannot fill the
machine on its own (O-proj and down: 32 weight tiles): token tile x cluster split-K x ring depth.  The production
heuristic (ops.pick_prefill_tile / pick_splitk) is printed next to the sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
C = ops.native(); C.init_kernels(0)
dev = "cuda"
H, F, QKV = 4096, 14336, 6144
shapes = {"qkv": (QKV, H, ops.EPI_PLAIN), "o": (H, H, ops.EPI_RESIDUAL), "down": (H, F, ops.EPI_RESIDUAL)}
ws = {k: [(torch.randn(n, kk, device=dev) * 0.02).bfloat16() for _ in range(3)] for k, (n, kk, _) in shapes.items()}

def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

for T in [int(v) for v in os.environ.get("B2B_PROBE_T", "512,256,1024").split(",")]:
    for name, (n, k, epi) in shapes.items():
        x = torch.randn(T, k, device=dev).bfloat16()
        res = torch.randn(T, n, device=dev).bfloat16() if epi == ops.EPI_RESIDUAL else None
        out = torch.empty(T, n, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * T * n * k
        def prod(i=0):
            ops.gemm(ws[name][i % 3], x, out=out, epi=epi, residual=res)
        bn_p, st_p, sk_p = ops.pick_prefill_tile(n, T, k)
        sk_p = sk_p or ops.pick_splitk(n, T, k, bn_p, epi, max(st_p, 0))
        us = timed(prod)
        print(f"T={T:5d} {name:5s} PRODUCTION bn={bn_p} stages={st_p} splitk={sk_p}: {us:7.1f} us {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
        for bn in (128, 256):
            if bn > T:
                continue
            for sk in (1, 2, 4):
                for st in (0, 3, 2):
                    cap = C.gemm_max_splitk(bn, epi, st)
                    if sk > cap or bn // sk < 4:
                        continue
                    def run(i=0):
                        ops.gemm(ws[name][i % 3], x, out=out, epi=epi, residual=res, bn=bn, splitk=sk, stages=st)
                    try:
                        us = timed(run)
                    except Exception as e:
                        print(f"T={T} {name} bn={bn} sk={sk} st={st}: FAILED {str(e)[:80]}")
                        continue
                    print(f"T={T:5d} {name:5s} bn={bn:3d} splitk={sk} stages={st}: {us:7.1f} us {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
