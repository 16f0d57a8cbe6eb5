"""Per-CTA timeline of the GEMM kernel (globaltimer): where do the fixed ~10 us go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops
C = ops.native(); C.init_kernels(0)
names = ["entry", "setup_done", "prefetch_issued", "first_full", "last_mma_commit", "tmem_full_seen", "epi_done", "exit"]
flush = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")
shapes = [(32, 4096, 4096, 1), (32, 4096, 4096, 4), (32, 28672, 4096, 1), (32, 4096, 14336, 4)]
if os.environ.get('TL_M'):
    shapes = [(int(m), n, k, sk) for m in os.environ['TL_M'].split(',') for (n, k, sk) in [(4096, 4096, 4), (6144, 4096, 2), (4096, 4096, 1)]]
for (m, n, k, sk) in shapes:
    w = (torch.randn(n, k, device="cuda") * 0.02).bfloat16(); x = torch.randn(m, k, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    nct = (n // 128) * sk
    dbg = torch.zeros(nct * 8, device="cuda", dtype=torch.int64)
    for it in range(3):
        flush.fill_(1); dbg.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(w, x, out=out, splitk=sk, dbg=dbg.data_ptr()); e1.record(); torch.cuda.synchronize()
    d = dbg.view(nct, 8).cpu().double()
    t0 = d[:, 0].min()
    print(f"\nshape m={m} n={n} k={k} splitk={sk} ctas={nct} event_time={e0.elapsed_time(e1)*1e3:.1f}us  span(entry_min..exit_max)={(d[:,7].max()-t0)/1e3:.1f}us")
    for i, nm in enumerate(names):
        col = d[:, i]
        valid = col > 0
        if valid.any():
            c = (col[valid] - t0) / 1e3
            print(f"  {nm:16s} min {c.min():7.2f}  median {c.median():7.2f}  max {c.max():7.2f} us")
