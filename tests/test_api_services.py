"""HTTP sidecar + services + CLI (intent of the reference's tests/test_api.py, extended to the
success paths it never exercises: /chat local, streaming NDJSON, mesh fallback)."""
import asyncio
import json

import pytest
from fastapi.testclient import TestClient

from bee2bee_b200 import api as api_mod
from bee2bee_b200.services import (BaseService, EmbeddedOllama, HFRemoteService, HFService, OllamaService,
                                   ServiceError, build_service)


@pytest.fixture()
def client(monkeypatch):
    monkeypatch.setenv("BEE2BEE_TRANSPORT", "inproc")
    monkeypatch.delenv("BEE2BEE_API_KEY", raising=False)
    api_mod.node = None
    with TestClient(api_mod.app) as c:       # runs the real lifespan: creates + starts a P2PNode
        yield c
    api_mod.node = None


def test_home_shape_and_uptime(client):
    d = client.get("/").json()
    assert d["status"] == "ok" and d["node_id"] == d["peer_id"] and d["node_id"].startswith("peer-")
    assert set(d) == {"status", "node_id", "peer_id", "region", "models", "services", "metrics"}
    assert set(d["metrics"]) == {"uptime", "pool_size", "status"} and d["metrics"]["status"] == "active"
    assert d["metrics"]["uptime"] >= 0


def test_auth_required_only_when_key_configured(client, monkeypatch):
    for route in ("/peers", "/providers"):
        assert client.get(route).status_code == 200
    monkeypatch.setenv("BEE2BEE_API_KEY", "s3cret")
    for route in ("/peers", "/providers", "/metrics", "/topology"):
        assert client.get(route).status_code == 401
        assert client.get(route, headers={"X-API-KEY": "wrong"}).status_code == 401
        assert client.get(route, headers={"X-API-KEY": "s3cret"}).status_code == 200
    assert client.post("/chat", json={"prompt": "x"}).status_code == 401
    assert client.post("/generate", json={"prompt": "x"}, headers={"X-API-KEY": "wrong"}).status_code == 401
    assert client.get("/").status_code == 200            # home stays open
    assert isinstance(client.get("/peers", headers={"X-API-KEY": "s3cret"}).json(), list)


def test_chat_local_service_buffered_and_streaming(client):
    svc = HFService("tiny-llama", 0.001, max_batch=4, max_seq_len=256, device="cpu")
    svc.load_sync()
    api_mod.node.local_services["hf"] = svc
    home = client.get("/").json()
    assert home["models"] == ["tiny-llama"] and home["services"]["hf"]["price_per_token"] == 0.001
    r = client.post("/generate", json={"prompt": "user: hi", "model": "tiny", "max_new_tokens": 5, "temperature": 0}).json()
    assert r["status"] == "ok" and r["rid"].startswith("local-") and r["text"].startswith("user: hi")
    assert r["metadata"]["engine"] == "coithub-local" and r["metadata"]["service"] == "hf"
    assert r["metadata"]["tokens"] == 5
    # substring model match in the other direction, /chat alias
    assert client.post("/chat", json={"prompt": "x", "model": "org/tiny-llama-chat", "max_new_tokens": 2}).json()["status"] == "ok"
    with client.stream("POST", "/generate", json={"prompt": "user: hi\nassistant:", "max_new_tokens": 6,
                                                  "stream": True}) as resp:
        assert resp.headers["content-type"].startswith("text/plain")
        lines = [json.loads(l) for l in resp.iter_lines() if l]
    assert lines[-1] == {"done": True} and all("text" in l for l in lines[:-1])
    # errors are HTTP 200 with status=error (reference contract)
    r = client.post("/generate", json={"prompt": "x", "model": "no-such-model"})
    assert r.status_code == 200 and r.json()["status"] == "error"
    # engine counters: JSON per service, and the Prometheus text exposition format
    m = client.get("/metrics").json()["hf"]
    assert m["requests"] >= 3 and m["tokens_generated"] >= 13 and m["healthy"] is True and m["timeouts"] == 0
    prom = client.get("/metrics", params={"format": "prometheus"})
    assert prom.headers["content-type"].startswith("text/plain")
    rows = dict(l.rsplit(" ", 1) for l in prom.text.strip().splitlines())
    assert float(rows['bee2bee_requests{service="hf"}']) == m["requests"]
    assert float(rows['bee2bee_healthy{service="hf"}']) == 1.0 and 'bee2bee_prefix_cache_hit_rate{service="hf"}' in rows
    svc.model.engine.stop()


def test_connect_route(client):
    r = client.get("/connect", params={"addr": "inproc://does-not-exist"}).json()
    assert r["status"] == "error"
    r = client.get("/connect", params={"addr": "coithub.org://join?network=n&model=m&hash=h"}).json()
    assert r["status"] == "error"


def test_hf_service_contract():
    svc = HFService("tiny-gpt2", 0.5, max_batch=2, max_seq_len=128, device="cpu")
    with pytest.raises(ServiceError, match="Model not loaded"):
        svc.execute({"prompt": "x"})
    svc.load_sync()
    assert svc.get_metadata()["models"] == ["tiny-gpt2"] and svc.get_metadata()["max_new_tokens"] == 2048
    with pytest.raises(ServiceError, match="Missing prompt"):
        svc.execute({})
    r = svc.execute({"prompt": "abc", "max_new_tokens": 4, "temperature": 0.0})
    assert set(r) == {"text", "tokens", "latency_ms", "price_per_token", "cost"}
    assert r["tokens"] == 4 and r["cost"] == 2.0 and r["text"].startswith("abc")
    assert svc.serves("tiny") and svc.serves("x/tiny-gpt2-y") and not svc.serves("llama") and svc.serves(None)
    # greedy is deterministic and temperature 0 is honoured (the reference turns 0 into 0.7)
    assert svc.execute({"prompt": "abc", "max_new_tokens": 4, "temperature": 0})["text"] == r["text"]
    with pytest.raises(ServiceError):
        HFService("definitely-not-a-model", 0).load_sync()
    svc.model.engine.stop()


def test_ollama_service_embedded_backend(monkeypatch):
    monkeypatch.setenv("OLLAMA_HOST", "embedded")
    svc = OllamaService("tiny-llama", max_batch=2, max_seq_len=128, device="cpu")
    assert svc.host == "embedded"                     # OLLAMA_HOST is honoured
    svc.load_sync()
    meta = svc.get_metadata()
    assert meta == {"models": ["tiny-llama", "tiny-llama:latest"], "price_per_token": 0.0, "backend": "ollama"}
    r = svc.execute({"prompt": "hello", "max_new_tokens": 5})
    assert set(r) == {"text", "tokens", "latency_ms", "price_per_token", "cost"} and r["tokens"] == 5 and r["cost"] == 0.0
    chunks = list(svc.execute_stream({"prompt": "hello", "max_new_tokens": 5}))
    assert chunks and all(isinstance(c, str) for c in chunks)
    for c in chunks:                                  # raw text, NOT NDJSON
        try:
            assert not isinstance(json.loads(c), dict)
        except ValueError:
            pass
    tags = EmbeddedOllama("gemma2:2b").tags()
    assert tags["models"][0]["name"] == "gemma2:2b"


def test_hf_remote_service_with_stub_client():
    class Stub:
        def __init__(self, model, token):
            self.model, self.token, self.calls = model, token, []

        def text_generation(self, prompt, **kw):
            self.calls.append(kw)
            if kw.get("stream"):
                return iter(["re", "mote"])
            return "generated text!"

    svc = HFRemoteService("meta-llama/Llama-2-7b-hf", token="tok", client_factory=Stub)
    with pytest.raises(ServiceError, match="not initialized"):
        svc.execute({"prompt": "x"})
    svc.load_sync()
    assert svc.get_metadata() == {"models": ["meta-llama/Llama-2-7b-hf"], "price_per_token": 0.005, "tag": "remote",
                                  "backend": "hf_remote"}
    r = svc.execute({"prompt": "x", "max_new_tokens": 9})
    assert r["text"] == "generated text!" and r["tokens"] == len("generated text!") // 4 and r["backend"] == "hf_remote"
    assert svc.client.calls[0]["max_new_tokens"] == 9
    lines = [json.loads(l) for l in svc.execute_stream({"prompt": "x"})]
    assert lines == [{"text": "re"}, {"text": "mote"}, {"done": True}]
    assert isinstance(build_service("hf_remote", "m", token="t"), HFRemoteService)
    with pytest.raises(ServiceError):
        build_service("nope", "m")


def test_async_adapters_do_not_block_the_loop():
    class Slow(BaseService):
        def get_metadata(self):
            return {"models": ["m"]}

        def execute(self, params):
            import time
            time.sleep(0.3)
            return {"text": "done"}

        def execute_stream(self, params):
            import time
            for i in range(3):
                time.sleep(0.1)
                yield f"c{i}"

    async def go():
        svc = Slow("hf")
        ticks = 0

        async def ticker():
            nonlocal ticks
            while True:
                await asyncio.sleep(0.01)
                ticks += 1

        t = asyncio.create_task(ticker())
        res = await svc.aexecute({})
        chunks = [c async for c in svc.aexecute_stream({})]
        t.cancel()
        assert res == {"text": "done"} and chunks == ["c0", "c1", "c2"]
        assert ticks > 20           # the loop kept running while the service worked

    asyncio.run(go())


def test_cli_surface(monkeypatch):
    from click.testing import CliRunner

    from bee2bee_b200.__main__ import cli

    r = CliRunner()
    out = r.invoke(cli, ["--help"]).output
    for verb in ("serve-ollama", "serve-hf", "serve-hf-remote", "register", "config", "topology", "bench"):
        assert verb in out
    h = r.invoke(cli, ["serve-hf", "--help"]).output
    assert "--model" in h and "--port" in h and "--region" in h and "--api-port" in h and "--pieces" in h
    h = r.invoke(cli, ["serve-ollama", "--help"]).output
    assert "--public-host" in h and "--host" in h
    assert "--token" in r.invoke(cli, ["serve-hf-remote", "--help"]).output
    res = r.invoke(cli, ["serve-hf-remote"])                  # token is required
    assert res.exit_code != 0 and "token" in res.output.lower()
    res = r.invoke(cli, ["register", "--region", "EU", "--network", "testnet"])
    assert res.exit_code == 0 and "Handshake OK" in res.output and "registry.json" in res.output
    from bee2bee_b200.registry import RegistryClient
    rows = RegistryClient.local_rows()
    assert len(rows) == 1 and list(rows.values())[0]["tag"] == "cli-testnet" and list(rows.values())[0]["models"] == ["system-test"]
    res = r.invoke(cli, ["register", "--region", "EU", "--node-url", "ws://127.0.0.1:9"])
    assert res.exit_code == 1 and "Handshake FAILED" in res.output
    res = r.invoke(cli, ["config", "bootstrap_url", "ws://1.2.3.4:5"])
    assert res.exit_code == 0
    assert json.loads(r.invoke(cli, ["config"]).output)["bootstrap_url"] == "ws://1.2.3.4:5"


def test_cli_topology_prints_the_piece_plan():
    from click.testing import CliRunner

    from bee2bee_b200.__main__ import cli

    r = CliRunner()
    assert "--supervised" in r.invoke(cli, ["serve-hf", "--help"]).output
    res = r.invoke(cli, ["topology", "--model", "llama-3-8b", "--pieces", "8"])
    assert res.exit_code == 0, res.output
    d = json.loads(res.output)
    assert [p["n_units"] for p in d["pieces"]] == [12, 12, 13, 13, 13, 13, 13, 7]      # thirds of a layer, lm_head on the last
    assert d["pieces"][0]["extras"] == ["embedding"] and d["pieces"][-1]["extras"] == ["lm_head", "sampler"]
    assert d["pieces"][2]["layers"] == [8, 12] and d["pieces"][2]["tail_gemm_of"].startswith("attention block")
    assert d["model"]["layers"] == 32 and "can_access_peer" in d
