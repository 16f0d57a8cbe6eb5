"""Separates fixed per-kernel cost from streaming cost: t(K) for fixed grids, isolated vs chained in a graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
C = ops.native(); C.init_kernels(0)
dev = "cuda"
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=8, chain=1):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / chain)
    ts.sort()
    return ts[len(ts) // 2]

m = 32
print("== isolated kernel time vs K (n=4096 rows -> 32 tiles), splitk in {1,4}")
for sk in (1, 4):
    for k in (512, 1024, 2048, 4096, 8192, 16384):
        w = (torch.randn(4096, k, device=dev) * 0.02).bfloat16(); x = torch.randn(m, k, device=dev).bfloat16()
        out = torch.empty(m, 4096, device=dev, dtype=torch.bfloat16)
        t = timed(lambda: ops.gemm(w, x, out=out, splitk=sk))
        print(f"n=4096 k={k:6d} splitk={sk}: {t:7.1f} us   {4096*k*2/t/1e3:6.0f} GB/s", flush=True)
print("== isolated vs N at K=4096 (tiles = N/128), no split")
for n in (4096, 8192, 16384, 28672, 57344, 131072):
    w = (torch.randn(n, 4096, device=dev) * 0.02).bfloat16(); x = torch.randn(m, 4096, device=dev).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    t = timed(lambda: ops.gemm(w, x, out=out, splitk=1))
    print(f"n={n:6d} k=4096 tiles={n//128:5d}: {t:7.1f} us   {n*4096*2/t/1e3:6.0f} GB/s", flush=True)
print("== chain of 16 distinct GEMMs (n=4096,k=4096,splitk=4) in one graph: per-kernel cost with/without PDL")
ws = [(torch.randn(4096, 4096, device=dev) * 0.02).bfloat16() for _ in range(16)]
xs = [torch.randn(m, 4096, device=dev).bfloat16() for _ in range(2)]
for pdl in (True, False):
    C.set_pdl(pdl)
    for sk in (1, 2, 4, 8):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ops.gemm(ws[0], xs[0], out=xs[1], splitk=sk); s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(16):
                    ops.gemm(ws[i], xs[i % 2], out=xs[(i + 1) % 2], splitk=sk)
            g.replay(); s.synchronize()
            ts = []
            for _ in range(6):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); g.replay(); e1.record(s); s.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / 16)
            ts.sort()
        print(f"chain pdl={int(pdl)} splitk={sk}: {ts[len(ts)//2]:6.1f} us/kernel ({4096*4096*2/ts[len(ts)//2]/1e3:5.0f} GB/s)", flush=True)
print("== empty-ish kernels in a graph: launch floor (rmsnorm on 32x4096)")
x = torch.randn(32, 4096, device=dev).bfloat16(); gmm = torch.ones(4096, device=dev).bfloat16(); o = torch.empty_like(x)
for pdl in (True, False):
    C.set_pdl(pdl)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.rmsnorm(x, gmm, out=o); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(64):
                ops.rmsnorm(x, gmm, out=o)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); s.synchronize()
    print(f"rmsnorm chain pdl={int(pdl)}: {e0.elapsed_time(e1)*1e3/64:5.2f} us/kernel")
