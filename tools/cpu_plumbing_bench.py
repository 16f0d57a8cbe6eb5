"""BASELINE config 1, measured: distilgpt2 split into 2 layer pieces hosted by two mesh peers on this machine (CPU),
hidden states hop over the loopback WebSocket runtime as binary `hidden_forward` frames, generation through the FastAPI
`POST /generate` route with a 16-token synthetic prompt.  Prints one JSON line with per-request latency, tokens/s and
the per-hop payload (the reference ships the same hop as a JSON nested list of fp32, node.py:270-277)."""
import argparse
import asyncio
import json
import os
os.environ.setdefault("B2B_ALLOW_RANDOM_WEIGHTS", "1")     # no checkpoints offline: random-init weights
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BEE2BEE_OFFLINE", "1")

import httpx  # noqa: E402
import uvicorn  # noqa: E402

from bee2bee_b200 import api as api_mod  # noqa: E402
from bee2bee_b200.models.config import resolve_config  # noqa: E402
from bee2bee_b200.p2p_runtime import P2PNode  # noqa: E402
from bee2bee_b200.parallel.cpu_pipeline import MeshPipelineService, PieceHost, piece_key  # noqa: E402
from bee2bee_b200.pieces import plan_pieces  # noqa: E402


async def main(a):
    cfg = resolve_config(a.model)
    head, tail = P2PNode(host="127.0.0.1", port=0), P2PNode(host="127.0.0.1", port=0)
    await head.start(); await tail.start()
    plan = plan_pieces(a.model, cfg.n_layers, 2, devices=["cpu"] * 2)
    tail.piece_hosts[piece_key(a.model, 1)] = PieceHost(a.model, 1, 2)
    tail.add_layer_piece(plan[1]); head.add_layer_piece(plan[0])
    await head.connect_bootstrap(tail.addr)
    while tail.peer_id not in head.peers:
        await asyncio.sleep(0.01)
    svc = MeshPipelineService(head, a.model, 2, [tail.peer_id])
    svc.bind_loop(asyncio.get_running_loop())
    await head.add_service(svc)
    api_mod.node = head
    port = a.port
    server = uvicorn.Server(uvicorn.Config(api_mod.app, host="127.0.0.1", port=port, log_level="warning"))
    task = asyncio.create_task(server.serve())
    await asyncio.sleep(0.5)
    prompt = " ".join(f"w{i}" for i in range(16))
    lats, toks = [], 0
    async with httpx.AsyncClient(timeout=300) as c:
        for i in range(a.warmup + a.requests):
            t0 = time.perf_counter()
            r = (await c.post(f"http://127.0.0.1:{port}/generate",
                              json={"prompt": prompt, "model": a.model, "max_new_tokens": a.max_new_tokens, "temperature": 0.7})).json()
            dt = time.perf_counter() - t0
            assert r.get("status") == "ok", r
            if i >= a.warmup:
                lats.append(dt)
                toks += int((r.get("metadata") or {}).get("tokens") or a.max_new_tokens)
    hop_bytes = cfg.hidden_size * 4                      # fp32 hidden state of one decode token per hop
    out = {"config": f"{a.model} x2 pieces, CPU, loopback ws, /generate, 16-token prompt", "requests": len(lats),
           "max_new_tokens": a.max_new_tokens, "latency_s_p50": statistics.median(lats), "latency_s_max": max(lats),
           "tokens_per_s": toks / sum(lats), "ms_per_token": 1e3 * sum(lats) / toks,
           "hop_payload_bytes_per_decode_token": hop_bytes, "hops_per_token": 2}
    print(json.dumps(out))
    server.should_exit = True
    await task
    await head.stop(); await tail.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="distilgpt2")
    ap.add_argument("--requests", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--max-new-tokens", type=int, default=16)
    ap.add_argument("--port", type=int, default=8017)
    asyncio.run(main(ap.parse_args()))
