"""Manual probe (parity: /root/reference/scripts/test_connection.py): hello -> peer_list ->
ping/pong against a live node.    python scripts/test_connection.py ws://127.0.0.1:4334"""
import asyncio
import json
import sys
import time

import websockets


async def main(addr: str):
    async with websockets.connect(addr, max_size=32 * 1024 * 1024) as ws:
        await ws.send(json.dumps({"type": "hello", "peer_id": "probe", "addr": None, "services": {}}))
        t0 = time.time()
        while time.time() - t0 < 3:
            m = json.loads(await asyncio.wait_for(ws.recv(), 3))
            print("<-", m["type"], {k: v for k, v in m.items() if k not in ("type", "metrics")})
            if m["type"] == "ping":
                await ws.send(json.dumps({"type": "pong", "ts": m["ts"]}))
                await ws.send(json.dumps({"type": "ping", "ts": time.time()}))
            if m["type"] == "pong":
                print(f"rtt {1000 * (time.time() - m['ts']):.2f} ms")
                return


if __name__ == "__main__":
    asyncio.run(main(sys.argv[1] if len(sys.argv) > 1 else "ws://127.0.0.1:4334"))
