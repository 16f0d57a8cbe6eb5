"""Multi-rank pipeline check (run under torchrun): N pieces over N GPUs with the fused NVLink handoff must produce
exactly the tokens of the single-GPU engine.

env: B2B_MODEL, B2B_STEPS, B2B_GROUPS, B2B_BATCH (sequences per group), B2B_PF_TOKENS (prefill chunk budget),
     B2B_PROMPTS = ramp (5 + 3 i tokens) | bigsmall (alternating long / 3-token prompts: many chunks of very different
     cost through the double-buffered prefill channel) | long (prompts of several chunks each),
     B2B_ENGINE=1 drives everything through Engine.generate (scheduler, fetch_window read-back) instead of the runner."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.parallel.dist import init_distributed, shutdown

model = os.environ.get("B2B_MODEL", "tiny-llama")
steps = int(os.environ.get("B2B_STEPS", "12"))
groups = int(os.environ.get("B2B_GROUPS", "0"))
B = int(os.environ.get("B2B_BATCH", "4"))
pf_tokens = int(os.environ.get("B2B_PF_TOKENS", "128"))
kind = os.environ.get("B2B_PROMPTS", "ramp")
quant = os.environ.get("B2B_QUANT", "bf16")
rank, world, local = init_distributed()
cfg = resolve_config(model)
groups = groups or world
total = B * groups
V = cfg.vocab_size


def plen(i):
    if kind == "bigsmall":
        return (pf_tokens - 4) if i % 2 == 0 else 3
    if kind == "long":
        return 2 * pf_tokens + 7 * i + 1
    return 5 + 3 * i


max_len = max(plen(i) for i in range(total)) + steps + 2
pages_per = (max_len + 63) // 64
max_seq = pages_per * 64
prompts = [[(13 * i + 7 * j + 5) % (V - 8) + 4 for j in range(plen(i))] for i in range(total)]

if os.environ.get("B2B_ENGINE") == "1":
    from bee2bee_b200.engine.core import Engine, SamplingParams
    eng = Engine(model, cfg=cfg, device=f"cuda:{local}", max_batch=total, groups=groups, max_seq_len=max_seq,
                 max_prefill_tokens=pf_tokens, decode_burst=max(1, steps // 3), rank=rank, world=world, quant=quant)
    half = total // 2
    sp = SamplingParams(max_new_tokens=steps + 1, temperature=0.0, top_p=1.0, repetition_penalty=1.0, ignore_eos=True)
    # two waves: the second prefill lands between decode bursts of the first (flags are never reset)
    reqs = [eng.submit(p, sp) for p in prompts[:half]]
    eng.step()
    reqs += [eng.submit(p, sp) for p in prompts[half:]]
    while not all(r.done.is_set() for r in reqs):
        eng.step()
    toks = [r.out_ids for r in reqs]
    if rank == 0:
        print("RESULT " + json.dumps({"world": world, "tokens": toks}), flush=True)
    eng.close()
else:
    r = GpuRunner(cfg, "", rank, world, torch.device(f"cuda:{local}"), max_batch=total, groups=groups, max_seq_len=max_seq,
                  max_prefill_tokens=pf_tokens, seed=0, quant=quant)
    seqs = [SeqInit(slot=i, prompt=prompts[i], pages=list(range(1 + pages_per * i, 1 + pages_per * (i + 1))),
                    temperature=0.0, top_p=1.0, repetition_penalty=1.0, seed=i) for i in range(total)]
    r.prefill(seqs)
    stall = float(os.environ.get("B2B_FAULT_STALL_S", "0"))
    if stall and rank == int(os.environ.get("B2B_FAULT_STALL_RANK", "1")):
        import time
        time.sleep(stall)                 # fault injection: this rank is late by `stall` seconds
    if stall:
        from bee2bee_b200.engine.runner import MeshStalled
        import time
        t0 = time.time()
        try:
            r.decode(steps // 2)
            r.sync()
            outcome = {"stalled": False}
        except MeshStalled as e:
            outcome = {"stalled": True, "after_s": round(time.time() - t0, 2), "error": str(e)[:80]}
        # the process survived: CUDA still works on this rank
        outcome["cuda_ok"] = bool(torch.ones(4, device=f"cuda:{local}").sum().item() == 4.0)
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, outcome)
        if rank == 0:
            print("RESULT " + json.dumps({"world": world, "tokens": gathered}), flush=True)
        r.close()
        shutdown()
        sys.exit(0)
    r.decode(steps // 2)
    r.sync()
    r.decode(steps - steps // 2)          # second burst: epochs are monotonic, nothing is re-armed in between
    win = r.fetch_window([0] * total, steps + 1)
    launches = [r.launches_per_decode_step()]          # kernels recorded per group step on each rank
    if world > 1:
        import torch.distributed as dist
        launches = [None] * world
        dist.all_gather_object(launches, r.launches_per_decode_step())
    if rank == 0:
        print("RESULT " + json.dumps({"world": world, "tokens": win[:total].tolist(), "chunks": r.pf_chunks,
                                      "launches": launches}), flush=True)
    r.close()
shutdown()
