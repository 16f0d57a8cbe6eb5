"""Manual probe: raw-WebSocket `gen_request` against a live node, buffered or streamed
(parity: /root/reference/scripts/test_generation.py and test_full_request.py).

    python -m bee2bee_b200 serve-hf --model tiny-llama --port 4334 --api-port 0 &
    python scripts/test_generation.py ws://127.0.0.1:4334 --prompt "user: hi" --stream
"""
import argparse
import asyncio
import json
import time
import uuid

import websockets


async def main(a):
    async with websockets.connect(a.addr, max_size=32 * 1024 * 1024) as ws:
        await ws.send(json.dumps({"type": "hello", "peer_id": f"probe-{uuid.uuid4().hex[:6]}", "addr": None, "services": {}}))
        rid = f"req-{uuid.uuid4().hex[:8]}"
        providers, sent, t0, text = {}, False, time.time(), []
        while time.time() - t0 < a.timeout:
            m = json.loads(await asyncio.wait_for(ws.recv(), a.timeout))
            kind = m.get("type")
            if kind == "hello":
                providers = m.get("services") or {}
                print("<- hello from", m.get("peer_id"), "services:", {k: v.get("models") for k, v in providers.items()})
                if not sent:
                    svc = a.svc or next(iter(providers), "hf")
                    model = a.model or (providers.get(svc, {}).get("models") or [None])[0]
                    await ws.send(json.dumps({"type": "gen_request", "rid": rid, "svc": svc, "model": model, "prompt": a.prompt,
                                              "max_new_tokens": a.max_new_tokens, "temperature": a.temperature, "stream": a.stream}))
                    sent, t0 = True, time.time()
                    print(f"-> gen_request rid={rid} svc={svc} model={model} stream={a.stream}")
            elif kind == "ping":
                await ws.send(json.dumps({"type": "pong", "ts": m["ts"]}))
            elif kind == "gen_chunk" and m.get("rid") == rid:
                text.append(m.get("text", ""))
                print(m.get("text", ""), end="", flush=True)
            elif kind in ("gen_success", "gen_result", "gen_error") and m.get("rid") == rid:
                dt = time.time() - t0
                if m.get("error"):
                    print(f"\n<- {kind}: ERROR {m['error']}")
                else:
                    out = m.get("text") or "".join(text)
                    print(f"\n<- {kind} in {dt * 1e3:.0f} ms: {out!r}  tokens={m.get('tokens')} latency_ms={m.get('latency_ms')}")
                return
        print("timed out")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("addr", nargs="?", default="ws://127.0.0.1:4334")
    ap.add_argument("--prompt", default="user: Say hello to the mesh.")
    ap.add_argument("--model", default=None)
    ap.add_argument("--svc", default=None)
    ap.add_argument("--max-new-tokens", type=int, default=32)
    ap.add_argument("--temperature", type=float, default=0.7)
    ap.add_argument("--stream", action="store_true")
    ap.add_argument("--timeout", type=float, default=60.0)
    asyncio.run(main(ap.parse_args()))
