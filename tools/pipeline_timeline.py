"""Per-rank timeline of the wavefront (torchrun): for each (step, group) the device time spent in the
group's graph (spin-wait on the upstream flag + compute) and the gaps between graphs."""
import os
os.environ.setdefault("B2B_ALLOW_RANDOM_WEIGHTS", "1")     # no checkpoints offline: random-init weights
os.environ["B2B_GRAPH_PER_GROUP"] = "1"          # one graph per group so that a group-stage can be bracketed by events (production: one graph per step)
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bee2bee_b200.engine.runner import GpuRunner, SeqInit
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.parallel.dist import init_distributed, shutdown

rank, world, local = init_distributed()
cfg = resolve_config(os.environ.get("B2B_MODEL", "llama-3-8b"))
B = int(os.environ.get("B2B_BATCH", "32")); groups = int(os.environ.get("B2B_GROUPS", str(world))); STEPS = 12
total = B * groups
r = GpuRunner(cfg, "", rank, world, torch.device(f"cuda:{local}"), max_batch=total, groups=groups, max_seq_len=512,
              max_prefill_tokens=512)
seqs = [SeqInit(slot=i, prompt=[(7 + 131 * i + 31 * j) % 100000 + 256 for j in range(16)], pages=[1 + 2 * i, 2 + 2 * i],
                temperature=0.7, seed=i) for i in range(total)]
r.prefill(seqs); r.decode(8); r.sync()
r.prepare_burst()
ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(groups)] for _ in range(STEPS)]
base = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier()
with torch.cuda.stream(r.stream):
    base.record(r.stream)
    for s in range(STEPS):
        for g in range(groups):
            ev[s][g][0].record(r.stream); r.graphs[g].replay(); ev[s][g][1].record(r.stream)
r.stream.synchronize(); dist.barrier()
dur = [[ev[s][g][0].elapsed_time(ev[s][g][1]) * 1e3 for g in range(groups)] for s in range(STEPS)]
start = [[base.elapsed_time(ev[s][g][0]) * 1e3 for g in range(groups)] for s in range(STEPS)]
end_total = base.elapsed_time(ev[-1][-1][1]) * 1e3
steady = dur[4:]
flat = [d for row in steady for d in row]
gaps = []
for s in range(4, STEPS):
    for g in range(groups):
        if g + 1 < groups:
            gaps.append(start[s][g + 1] - (start[s][g] + dur[s][g]))
info = {"rank": rank, "units": list(r.units), "launches": r.launches_per_decode_step(), "graph_us_mean": sum(flat) / len(flat), "graph_us_min": min(flat),
        "graph_us_max": max(flat), "gap_us_mean": sum(gaps) / max(1, len(gaps)), "per_step_us": (start[-1][0] - start[4][0]) / (STEPS - 5),
        "total_us": end_total}
out = [None] * world
dist.all_gather_object(out, info)
if rank == 0:
    for o in out:
        print("TL " + json.dumps(o), flush=True)
r.close(); shutdown()
