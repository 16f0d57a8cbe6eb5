"""Manual probe: a client-side P2PNode joins a mesh, lists providers, picks one by (price, latency) and requests a
generation through `request_generation` (parity: /root/reference/scripts/debug_p2p_request.py, debug_generation.py).

    python scripts/debug_p2p_request.py ws://127.0.0.1:4334 --model tiny-llama --prompt "user: hi"
"""
import argparse
import asyncio
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bee2bee_b200.p2p_runtime import P2PNode  # noqa: E402


async def main(a):
    node = P2PNode(host="127.0.0.1", port=0, region="Probe")
    await node.start()
    try:
        ok = await node.connect_bootstrap(a.bootstrap)
        print("connected:", ok, "| my addr", node.addr)
        t0 = time.time()
        while not node.list_providers() and time.time() - t0 < 5:
            await asyncio.sleep(0.05)
        for p in node.list_providers():
            print("provider", p["peer_id"], p["addr"], "models", p["models"], "price", p["price_per_token"], "rtt_ms", p["latency_ms"])
        picked = node.pick_provider(a.model) if a.model else None
        pid = picked[0] if picked else next((p["peer_id"] for p in node.list_providers() if p["peer_id"] != node.peer_id), None)
        if pid is None:
            print("no provider found")
            return
        chunks = []
        t0 = time.time()
        res = await node.request_generation(pid, a.prompt, max_new_tokens=a.max_new_tokens, model_name=a.model,
                                            temperature=a.temperature, on_chunk=(lambda c: (chunks.append(c), print(c, end="", flush=True)))
                                            if a.stream else None, timeout=a.timeout)
        print(f"\nresult from {pid} in {(time.time() - t0) * 1e3:.0f} ms:", {k: v for k, v in res.items() if k != "text"})
        print("text:", repr(res.get("text") or "".join(chunks)))
    finally:
        await node.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("bootstrap", nargs="?", default="ws://127.0.0.1:4334")
    ap.add_argument("--model", default=None)
    ap.add_argument("--prompt", default="user: Hello from a peer.")
    ap.add_argument("--max-new-tokens", type=int, default=32)
    ap.add_argument("--temperature", type=float, default=0.7)
    ap.add_argument("--stream", action="store_true")
    ap.add_argument("--timeout", type=float, default=60.0)
    asyncio.run(main(ap.parse_args()))
