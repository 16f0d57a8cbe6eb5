"""``torch.distributed`` plumbing: one process per GPU, NCCL over NVLink for the control
collectives (barriers, handle exchange, result broadcast); gloo on CPU-only hosts.  The
token path itself never goes through NCCL (see ``parallel.mesh``)."""
from __future__ import annotations

import datetime
import os
from typing import Optional, Tuple

import torch


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def init_distributed(backend: Optional[str] = None, timeout_s: int = 600, eager: bool = True) -> Tuple[int, int, int]:
    """Initialise from torchrun-style env vars. Returns (rank, world, local_rank).  ``eager=False`` lets NCCL create its
    communicators lazily: unbatched send / recv then run on dedicated per-pair communicators instead of being serialised
    with every other op of the group (what the constructed NCCL pipeline arm wants)."""
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world <= 1:
        return 0, 1, 0
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl" and eager:
            kw["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=timeout_s), **kw)
    return rank, world, local


def max_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shutdown() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()
