"""NVLink handoff measurements between rank 0 and rank 1 (run under torchrun, >= 2 ranks):
  1. device-timed flag round trip (the mesh's replacement for the reference's ping/pong RTT),
  2. peer write bandwidth: kernel stores into IPC-mapped memory and cudaMemcpyPeerAsync,
  3. the fused piece-tail GEMM (down-proj + residual) with its epilogue storing locally vs into the peer."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bee2bee_b200 import ops
from bee2bee_b200.parallel.dist import init_distributed, max_over_ranks, shutdown

rank, world, local = init_distributed()
dev = torch.device(f"cuda:{local}")
C = ops.native(); C.init_kernels(local)
H, F = 4096, 14336
NB = 256 << 20
# symmetric buffers: [0] flags (4 KiB), [1] payload
flags_p, data_p = C.peer_alloc(4096), C.peer_alloc(NB)
handles = [None] * world
dist.all_gather_object(handles, {"flags": C.ipc_export(flags_p), "data": C.ipc_export(data_p)})
peer = (rank + 1) % world if rank < 2 else rank
out = {}
if rank < 2:
    other = 1 - rank
    pf, pd = C.ipc_import(handles[other]["flags"]), C.ipc_import(handles[other]["data"])
dist.barrier()
res = {}
if rank < 2:
    # ---- 1. flag ping-pong: word 0 = my in-flag (peer writes), word 1 = my epoch (consumed), word 2 = my out-epoch
    ITERS = 2000
    s = torch.cuda.Stream()
    my_flag, my_epoch, my_out_epoch = flags_p, flags_p + 4, flags_p + 8
    peer_flag = pf
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(100):
                if rank == 0:
                    C.flag_signal(peer_flag, my_out_epoch, 0, 0)
                    C.flag_wait(my_flag, my_epoch, 1); C.flag_signal(0, 0, my_epoch, 0)
                else:
                    C.flag_wait(my_flag, my_epoch, 1); C.flag_signal(0, 0, my_epoch, 0)
                    C.flag_signal(peer_flag, my_out_epoch, 0, 0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        g.replay(); e0.record(s)
        for _ in range(ITERS // 100):
            g.replay()
        e1.record(s)
    s.synchronize()
    res["flag_round_trip_us"] = e0.elapsed_time(e1) * 1e3 / ITERS
    dist.barrier()
    # ---- 2. peer write bandwidth
    src = torch.empty(NB, dtype=torch.uint8, device=dev).fill_(7)
    dst_peer = C.tensor_from_ptr(pd, [NB], "u8", local)
    for name, fn in (("kernel_store_GBps", lambda: dst_peer.copy_(src)),
                     ("memcpy_peer_GBps", lambda: C.memcpy_peer(pd, other, src.data_ptr(), local, NB))):
        if rank == 0:
            fn(); torch.cuda.synchronize()
            e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
            res[name] = 5 * NB / (e0.elapsed_time(e1) / 1e3) / 1e9
        dist.barrier()
    # ---- 3. fused tail GEMM: local vs peer epilogue stores
    w = (torch.randn(H, F, device=dev) * 0.02).bfloat16()
    for T in (32, 512, 4096):
        x = torch.randn(T, F, device=dev).bfloat16(); r = torch.randn(T, H, device=dev).bfloat16()
        out_local = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        if rank == 0:
            for name, kw in (("local", dict(out=out_local)), ("peer", dict(out_ptr=pd, ld_out=H))):
                ts = []
                for it in range(6):
                    flush.fill_(it)
                    e0.record(); ops.gemm(w, x, epi=ops.EPI_RESIDUAL, residual=r, **kw); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                res[f"tail_gemm_T{T}_{name}_us"] = sorted(ts)[len(ts) // 2]
            pay = T * H * 2
            res[f"tail_gemm_T{T}_payload_MB"] = pay / 1e6
            res[f"tail_gemm_T{T}_peer_GBps_lower_bound"] = pay / res[f"tail_gemm_T{T}_peer_us"] / 1e3
        dist.barrier()
        if rank == 1 and T == 4096:      # the consumer really received the tiles
            got = C.tensor_from_ptr(data_p, [T, H], "bf16", local).float()
            res["peer_payload_finite"] = bool(torch.isfinite(got).all()) and float(got.abs().sum()) > 0
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    merged = {}
    for g_ in gathered:
        for k, v in g_.items():
            merged.setdefault(k, v)
    print("HANDOFF " + json.dumps(merged), flush=True)
shutdown()
