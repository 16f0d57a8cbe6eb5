"""Mesh runtime over the in-process transport (N peers in one process, no ports) and over
loopback WebSockets: hello / peer_list gossip, ping/pong RTT, service announce, local +
streamed + relayed generation, provider selection, failure detection, byte-piece exchange."""
import asyncio
import json

import pytest

from bee2bee_b200 import protocol as P
from bee2bee_b200.p2p_runtime import P2PNode
from bee2bee_b200.services import BaseService, ServiceError


class EchoService(BaseService):
    def __init__(self, name="hf", models=("echo-model",), price=0.0, tag=None, delay=0.0):
        super().__init__(name)
        self.models, self.price, self.tag, self.delay, self.calls = list(models), price, tag, delay, 0

    def get_metadata(self):
        m = {"models": self.models, "price_per_token": self.price, "max_new_tokens": 64}
        if self.tag:
            m["tag"] = self.tag
        return m

    def execute(self, params):
        import time
        self.calls += 1
        time.sleep(self.delay)
        if params["prompt"] == "boom":
            raise ServiceError("kaput")
        return {"text": params["prompt"][::-1], "tokens": len(params["prompt"]), "latency_ms": 1,
                "price_per_token": self.price, "cost": 0.0, "max_new": params.get("max_new_tokens")}

    def execute_stream(self, params):
        for ch in params["prompt"]:
            yield json.dumps({"text": ch}) + "\n"
        yield json.dumps({"done": True}) + "\n"


async def mesh(n, transport="inproc"):
    nodes = [P2PNode(host="127.0.0.1", port=0, transport=transport, name=f"n{i}-{id(object())}", health_interval=0.2,
                     pong_timeout=1.0) for i in range(n)]
    for nd in nodes:
        await nd.start()
    return nodes


async def settle(cond, timeout=5.0):
    t0 = asyncio.get_running_loop().time()
    while not cond():
        if asyncio.get_running_loop().time() - t0 > timeout:
            raise AssertionError("condition not reached")
        await asyncio.sleep(0.01)


def run(coro):
    return asyncio.run(coro)


def test_hello_gossip_full_mesh_and_rtt():
    async def go():
        a, b, c = await mesh(3)
        try:
            await b.connect_bootstrap(a.addr)
            await c.connect_bootstrap(a.addr)          # c learns about b through a's peer_list
            await settle(lambda: all(len(n.peers) == 2 and all(p.get("hello_seen") for p in n.peers.values())
                                     for n in (a, b, c)))
            assert set(a.peers) == {b.peer_id, c.peer_id} and set(c.peers) == {a.peer_id, b.peer_id}
            await settle(lambda: all(p.get("last_pong_at") for n in (a, b, c) for p in n.peers.values()))
            assert all(0 <= p["last_pong_ms"] < 1000 for p in a.peers.values())
            assert a.peers[b.peer_id]["metrics"] is not None           # metrics piggy-backed on hello/ping
            topo = a.mesh_topology()
            assert set(topo) == {a.peer_id, b.peer_id, c.peer_id}
        finally:
            for n in (a, b, c):
                await n.stop()

    run(go())


def test_generation_local_stream_relay_and_errors():
    async def go():
        a, b, c = await mesh(3)
        try:
            svc = EchoService()
            await b.add_service(svc)
            await a.connect_bootstrap(b.addr)
            await c.connect_bootstrap(a.addr)
            await settle(lambda: b.peer_id in a.providers and b.peer_id in c.providers and len(c.peers) == 2)
            # python requester resolves on gen_success (reference never does: SURVEY R8)
            res = await a.request_generation(b.peer_id, "hello", 7, "echo-model", timeout=5)
            assert res["text"] == "olleh" and res["max_new"] == 7
            # streaming: gen_chunk per delta
            chunks = []
            res = await a.request_generation(b.peer_id, "abc", 4, "echo-model", on_chunk=chunks.append, timeout=5)
            assert chunks == ["a", "b", "c"]
            # provider-side error -> gen_error -> ServiceError at the requester
            with pytest.raises(ServiceError, match="local_error: kaput"):
                await a.request_generation(b.peer_id, "boom", 4, "echo-model", timeout=5)
            # self request shortcut
            assert (await b.request_generation(b.peer_id, "xy", 4, "echo-model"))["text"] == "yx"
            # relay: ask a peer that does not serve the model; it forwards to the provider (gen_result)
            await c._drop_peer(b.peer_id)           # c only knows b through a now
            c.providers.pop(b.peer_id, None)
            conn = c._conn_of(a.peer_id)
            fut = asyncio.get_running_loop().create_future()
            c._pending_requests["r1"] = fut
            await c._send(conn, P.gen_request("r1", "relay", model="echo-model", max_new_tokens=3))
            out = await asyncio.wait_for(fut, 5)
            assert out["text"] == "yaler"
            fut = asyncio.get_running_loop().create_future()
            c._pending_requests["r2"] = fut
            await c._send(conn, P.gen_request("r2", "x", model="nobody-serves-this"))
            out = await asyncio.wait_for(fut, 5)
            assert out["error"] == "consensus_deadlock: no_node_available"
        finally:
            for n in (a, b, c):
                await n.stop()

    run(go())


def test_provider_listing_and_selection_policy():
    async def go():
        a, b, c, d = await mesh(4)
        try:
            await b.add_service(EchoService(price=0.002))
            await c.add_service(EchoService(price=0.001, tag="remote"))
            await d.add_service(EchoService(name="ollama", models=("echo-model", "echo-model:latest"), price=0.001))
            for n in (b, c, d):
                await a.connect_bootstrap(n.addr)
            await settle(lambda: len(a.providers) == 3 and all(v.get("_latency") is not None for v in a.providers.values()))
            a.providers[c.peer_id]["_latency"] = 9.0
            a.providers[d.peer_id]["_latency"] = 3.0
            pid, meta = a.pick_provider("echo-model")
            assert pid == d.peer_id and meta["_svc_name"] == "ollama"        # cheapest, then lowest latency
            a.providers[d.peer_id]["health"] = "degraded"
            assert a.pick_provider("echo-model")[0] == c.peer_id
            assert a.pick_provider("unknown") is None
            rows = {r["peer_id"]: r for r in a.list_providers()}
            assert rows[c.peer_id]["tag"] == "remote" and rows[b.peer_id]["price_per_token"] == 0.002
            assert rows[d.peer_id]["models"] == ["echo-model", "echo-model:latest"]
        finally:
            for n in (a, b, c, d):
                await n.stop()

    run(go())


def test_failure_detection_and_pending_request_failover():
    async def go():
        a, b = await mesh(2)
        try:
            await b.add_service(EchoService(delay=0.5))
            await a.connect_bootstrap(b.addr)
            await settle(lambda: b.peer_id in a.providers)
            task = asyncio.create_task(a.request_generation(b.peer_id, "slow", 4, "echo-model", timeout=10))
            await asyncio.sleep(0.1)
            await b.stop()                                  # provider dies mid-request
            with pytest.raises(ServiceError, match="relay_link_failure"):
                await asyncio.wait_for(task, 5)             # fails fast instead of the 300 s timeout
            await settle(lambda: b.peer_id not in a.peers and b.peer_id not in a.providers)
        finally:
            await a.stop()

    run(go())


def test_silent_peer_is_dropped_by_pong_timeout():
    async def go():
        a, b = await mesh(2)
        try:
            await a.connect_bootstrap(b.addr)
            await settle(lambda: b.peer_id in a.peers and a.peers[b.peer_id].get("last_pong_at"))
            b._handlers[P.PING] = lambda conn, data: asyncio.sleep(0)       # b stops answering pings
            a._bootstrap_addrs.clear()      # no automatic re-dial: the drop must stay observable
            await settle(lambda: b.peer_id not in a.peers, timeout=6)
        finally:
            await a.stop(); await b.stop()

    run(go())


def test_byte_piece_exchange_verifies_hashes():
    async def go():
        a, b = await mesh(2)
        try:
            await a.connect_bootstrap(b.addr)
            await settle(lambda: b.peer_id in a.peers and a.peers[b.peer_id].get("hello_seen"))
            blob = bytes(range(256)) * 5000
            h = await b.publish_blob(blob, piece_size=100_000)
            assert await a.fetch_blob(b.peer_id, h, timeout=10) == blob
            with pytest.raises(ServiceError, match="unknown_content"):
                await a.fetch_blob(b.peer_id, "00" * 32, timeout=5)
            b.pieces[h]["chunks"][3] = b"corrupt"                             # tampered piece is rejected
            with pytest.raises(ValueError, match="hash_mismatch_at_3"):
                await a.fetch_blob(b.peer_id, h, timeout=10)
        finally:
            await a.stop(); await b.stop()

    run(go())


def test_websocket_transport_wire_compat():
    """A raw websockets client speaking the reference's frames (like app/api/bridge.js)."""
    async def go():
        import websockets

        (node,) = await mesh(1, transport="ws")
        try:
            await node.add_service(EchoService())
            assert node.addr.startswith("ws://127.0.0.1:")
            async with websockets.connect(node.addr) as ws:
                await ws.send(json.dumps({"type": "hello", "peer_id": "js-bridge", "addr": None, "services": {}}))
                got = {}
                while "hello" not in got or "ping" not in got:
                    m = json.loads(await asyncio.wait_for(ws.recv(), 5))
                    got[m["type"]] = m
                assert got["hello"]["services"]["hf"]["models"] == ["echo-model"]
                assert set(got["hello"]) >= {"peer_id", "addr", "region", "metrics", "services", "api_port", "public_ip"}
                await ws.send(json.dumps({"type": "ping", "ts": got["ping"]["ts"]}))
                await ws.send(json.dumps({"type": "gen_request", "task_id": "t1", "model": "echo-model", "prompt": "hi!",
                                          "stream": True}))
                texts, final = [], None
                while final is None:
                    m = json.loads(await asyncio.wait_for(ws.recv(), 5))
                    if m["type"] == "gen_chunk":
                        texts.append(m["text"])
                    elif m["type"] == "gen_success":
                        final = m
                assert texts == ["h", "i", "!"] and final["rid"] == "t1"
        finally:
            await node.stop()

    run(go())


def test_unhealthy_service_is_announced_and_sorted_last():
    """a provider whose GPU mesh aborted (engine.broken -> service metadata healthy = False) tells the mesh at
    the next health tick; every node's pick_provider then prefers the healthy replica even if it is more expensive
    (the reference can only drop peers whose socket closed, p2p_runtime.py:396-410)."""
    async def go():
        a, b, c = await mesh(3)
        try:
            class Flaky(EchoService):
                ok = True

                def get_metadata(self):
                    return {**super().get_metadata(), "healthy": self.ok}

            cheap = Flaky(price=0.0)
            await a.add_service(cheap)
            await b.add_service(EchoService(price=1.0))
            await b.connect_bootstrap(a.addr)
            await c.connect_bootstrap(a.addr)
            await settle(lambda: len(c.providers) >= 2 and len(a.providers) >= 2)
            assert c.pick_provider("echo-model")[0] == a.peer_id           # cheapest wins while healthy
            cheap.ok = False                                               # "mesh aborted"
            await settle(lambda: c.providers.get(a.peer_id, {}).get("hf", {}).get("healthy") is False, timeout=5.0)
            assert c.pick_provider("echo-model")[0] == b.peer_id
            assert a.pick_provider("echo-model")[0] == b.peer_id and a.providers[a.peer_id]["health"] == "degraded"
            cheap.ok = True                                                # supervisor restarted the mesh
            await settle(lambda: c.providers[a.peer_id]["hf"].get("healthy") is True, timeout=5.0)
            assert c.pick_provider("echo-model")[0] == a.peer_id
        finally:
            for n in (a, b, c):
                await n.stop()

    run(go())
