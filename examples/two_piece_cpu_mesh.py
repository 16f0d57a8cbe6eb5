"""BASELINE.json config 1 in one file: distilgpt2 split into two layer pieces hosted by two mesh
peers on this machine (CPU), hidden states hop over the loopback p2p runtime, generation through
the FastAPI `/generate` route."""
import os
os.environ.setdefault("B2B_ALLOW_RANDOM_WEIGHTS", "1")     # no checkpoints offline: random-init weights
import asyncio

import httpx
import uvicorn

from bee2bee_b200 import api as api_mod
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.p2p_runtime import P2PNode
from bee2bee_b200.parallel.cpu_pipeline import MeshPipelineService, PieceHost, piece_key
from bee2bee_b200.pieces import plan_pieces

MODEL, PIECES, API_PORT = "distilgpt2", 2, 8011


async def main():
    head, tail = P2PNode(host="127.0.0.1", port=0), P2PNode(host="127.0.0.1", port=0)
    await head.start(); await tail.start()
    plan = plan_pieces(MODEL, resolve_config(MODEL).n_layers, PIECES, devices=["cpu"] * PIECES)
    tail.piece_hosts[piece_key(MODEL, 1)] = PieceHost(MODEL, 1, PIECES)
    tail.add_layer_piece(plan[1]); head.add_layer_piece(plan[0])
    await head.connect_bootstrap(tail.addr)
    while tail.peer_id not in head.peers:
        await asyncio.sleep(0.01)
    svc = MeshPipelineService(head, MODEL, PIECES, [tail.peer_id])
    svc.bind_loop(asyncio.get_running_loop())
    await head.add_service(svc)
    api_mod.node = head
    server = uvicorn.Server(uvicorn.Config(api_mod.app, host="127.0.0.1", port=API_PORT, log_level="warning"))
    task = asyncio.create_task(server.serve())
    await asyncio.sleep(0.5)
    async with httpx.AsyncClient() as c:
        r = await c.post(f"http://127.0.0.1:{API_PORT}/generate", timeout=120,
                         json={"prompt": "sixteen synthetic tokens..", "model": MODEL, "max_new_tokens": 16})
        print(r.json())
        print("topology:", (await c.get(f"http://127.0.0.1:{API_PORT}/topology")).json())
    server.should_exit = True
    await task
    await head.stop(); await tail.stop()


if __name__ == "__main__":
    asyncio.run(main())
