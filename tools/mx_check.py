"""First-light check of the MX (block-scaled) fp8 GEMM against exact math on the quantised operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200 import ops

def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)

for (m, n, k, splitk) in [(1, 256, 256, 1), (33, 256, 512, 1), (128, 256, 384, 1), (200, 384, 1024, 2), (32, 4096, 4096, 4)]:
    w, x = bf(n, k, scale=0.05, seed=1), bf(m, k, scale=2.0, seed=2)
    # make block magnitudes vary a lot so that wrong scale-factor addressing is visible
    x = (x.float() * torch.exp2(torch.randint(-6, 6, (m, k // 32), device="cuda").float()).repeat_interleave(32, 1)).to(torch.bfloat16)
    w = (w.float() * torch.exp2(torch.randint(-4, 4, (n, k // 32), device="cuda").float()).repeat_interleave(32, 1)).to(torch.bfloat16)
    wq, sfa = ops.quantize_weight_mxfp8(w)
    bn = ops.pick_bn_mx(m)
    xq, sfb = ops.quant_mxfp8_rows(x, bn)
    torch.cuda.synchronize()
    xs, ws = ops.mx_unchunk(sfb, m, k, bn), ops.mx_unchunk(sfa, n, k, 128)
    xd, wd = ops.mx_dequant(xq, xs), ops.mx_dequant(wq, ws)
    qerr = ((xd - x.float()).abs() / (x.float().abs() + 1e-3)).max().item()
    out = ops.gemm(wq, xq, sfa=sfa, sfb=sfb, splitk=splitk, bn=bn)
    torch.cuda.synchronize()
    ref_q = xd @ wd.t()
    ref = x.float() @ w.float().t()
    e1 = (out.float() - ref_q).abs().max().item() / ref_q.abs().max().item()
    e2 = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"m={m} n={n} k={k} splitk={splitk} bn={bn}: quant rel err {qerr:.3f}; gemm vs exact-on-quantised {e1:.5f}; vs bf16 math {e2:.4f}", flush=True)
