"""Web gateway (`/api/p2p/*`): route shapes of the reference's Express app (app/api/index.js:16-216) served by a
mesh-client P2PNode; generation goes through the mesh `gen_request` path and is streamed back as text."""
import asyncio
import json
import threading

import pytest
from fastapi.testclient import TestClient

from bee2bee_b200.gateway import MeshBridge, MetricsStore, create_app
from bee2bee_b200.p2p import generate_join_link
from bee2bee_b200.p2p_runtime import P2PNode
from bee2bee_b200.services import BaseService


class Echo(BaseService):
    def __init__(self):
        super().__init__("hf")

    def get_metadata(self):
        return {"models": ["echo-model"], "price_per_token": 0.0, "max_new_tokens": 64}

    def execute(self, params):
        return {"text": params["prompt"].upper(), "tokens": 1, "latency_ms": 1, "price_per_token": 0.0, "cost": 0.0}

    def execute_stream(self, params):
        for ch in params["prompt"].upper():
            yield json.dumps({"text": ch}) + "\n"
        yield json.dumps({"done": True}) + "\n"


class ProviderThread:
    """a provider node living on its own event loop (the TestClient drives the gateway's loop)"""

    def __init__(self):
        self.loop = asyncio.new_event_loop()
        self.node = None
        self.ready = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        assert self.ready.wait(10)

    def _run(self):
        asyncio.set_event_loop(self.loop)

        async def up():
            self.node = P2PNode(host="127.0.0.1", port=0, transport="ws", region="Test-Region")
            await self.node.start()
            await self.node.add_service(Echo())
            self.ready.set()

        self.loop.run_until_complete(up())
        self.loop.run_forever()

    def stop(self):
        fut = asyncio.run_coroutine_threadsafe(self.node.stop(), self.loop)
        fut.result(10)
        self.loop.call_soon_threadsafe(self.loop.stop)
        self.thread.join(5)


@pytest.fixture()
def gateway(tmp_path, monkeypatch):
    monkeypatch.setenv("BEE2BEE_OFFLINE", "1")
    prov = ProviderThread()
    app = create_app(MeshBridge(transport="ws"), MetricsStore(str(tmp_path / "m.json")))
    with TestClient(app) as c:
        yield c, prov
    prov.stop()


def test_register_status_generate_metrics(gateway):
    c, prov = gateway
    # idle before any peer is known
    st = c.get("/api/p2p/status").json()
    assert st["status"] == "idle" and st["connected"] is False and st["mesh"] == {} and st["mode"]
    assert c.post("/api/p2p/register", json={}).status_code == 400
    # join link -> dial the provider's bootstrap address
    link = generate_join_link("connectit", "echo-model", "ab" * 32, [prov.node.addr])
    r = c.post("/api/p2p/register", json={"link": link}).json()
    assert r["status"] == "registered" and r["model"] == "echo-model" and r["connected"] is True
    assert r["activeNode"] == prov.node.addr
    import time
    for _ in range(100):                    # the provider's hello (region, services) arrives asynchronously
        st = c.get("/api/p2p/status").json()
        if "Test-Region" in st["mesh"] and st["mesh"]["Test-Region"][0]["models"]:
            break
        time.sleep(0.05)
    assert st["status"] == "active" and st["poolSize"] >= 1
    node_rows = st["mesh"]["Test-Region"]
    assert node_rows[0]["peer_id"] == prov.node.peer_id and node_rows[0]["models"] == ["echo-model"]
    # generation: streamed text pass-through over the mesh gen_request path (both body shapes)
    r = c.post("/api/p2p/generate", json={"prompt": "hello mesh", "model": "echo-model", "max_tokens": 16})
    assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
    assert r.text.strip() == "HELLO MESH"
    r = c.post("/api/p2p/generate", json={"task": {"prompt": "abc", "model": "default"}})
    assert r.text.strip() == "ABC"
    assert c.post("/api/p2p/generate", json={"model": "x"}).status_code == 400
    # errors after the first byte are appended to the stream, HTTP status stays 200
    r = c.post("/api/p2p/generate", json={"prompt": "x", "model": "echo-model", "targetNode": "127.0.0.1:9"})
    assert r.text.strip() == "X"                    # dead direct target -> falls back to the mesh
    # token estimate (chars / 4) landed in the metrics store
    m = c.get("/api/p2p/global_metrics").json()
    assert m["chats"] == 3 and m["tokens"] == 3 + 1 + 1
    assert c.post("/api/p2p/global_metrics", json={"tokens": 10}).json() == {"success": True}
    assert c.post("/api/p2p/global_metrics", json={"tokens": 0}).json() == {"success": False}
    assert c.get("/api/p2p/global_metrics").json()["tokens"] == 15
    # discovery action + 404 shape + the built-in page
    assert c.post("/api/p2p/status", json={"action": "discover_peer", "peer": {"addr": prov.node.addr}}).json() == {
        "status": "discovery_initiated"}
    nf = c.get("/api/p2p/nope")
    assert nf.status_code == 404 and "not found" in nf.json()["error"]
    assert "mesh console" in c.get("/").text


def test_generate_without_any_node_reports_error_in_stream(tmp_path, monkeypatch):
    monkeypatch.setenv("BEE2BEE_OFFLINE", "1")
    with TestClient(create_app(MeshBridge(transport="inproc"), MetricsStore(str(tmp_path / "m.json")))) as c:
        r = c.post("/api/p2p/generate", json={"prompt": "hi"})
        assert r.status_code == 200 and "[Error]: no_node_available" in r.text
        assert c.get("/api/p2p/global_metrics").json() == {"visits": 0, "chats": 0, "tokens": 0}


def test_web_console_page_is_served(gateway):
    """C21: landing / quick-register / dashboard / chat console; same /api/p2p/* routes as the reference's SPA."""
    client = gateway[0] if isinstance(gateway, (tuple, list)) else gateway
    page = client.get("/").text
    for needle in ("parseJoinLink", "/api/p2p/register", "/api/p2p/generate", "/api/p2p/status", "/api/p2p/global_metrics",
                   "setInterval(refresh,15000)", "user", "assistant:", "URLSearchParams(location.search)"):
        assert needle in page, needle
