"""Web gateway: the `/api/p2p/*` HTTP surface the reference ships as an Express app + WebSocket bridge
(`app/api/index.js:16-216`, `app/api/bridge.js:20-426`), re-built as a FastAPI app on top of an in-process
:class:`~bee2bee_b200.p2p_runtime.P2PNode` acting as a pure mesh *client*.

Routes (same paths, verbs and JSON keys, so the reference's React client keeps working against it):

* ``POST /api/p2p/register``        ``{link}`` -> parses the join link, dials its bootstrap addresses
* ``POST /api/p2p/generate``        ``{prompt | task{prompt,model,targetNode}, model, max_tokens, temperature}`` ->
  streamed ``text/event-stream`` pass-through of the provider's text (direct ``POST <node>/generate`` first, mesh
  ``gen_request`` second); errors after the first byte are appended as ``\\n\\n[Error]: ...``
* ``GET|POST /api/p2p/status``      pool / mesh-by-region telemetry, optional ``?target=host:port`` probe;
  ``POST {action:"discover_peer", peer:{addr}}`` dials a peer
* ``GET|POST /api/p2p/global_metrics``  token / chat counters

Differences by design: no Supabase dependency (counters live in ``$BEE2BEE_HOME/gateway_metrics.json``; a configured
registry still receives node rows through ``RegistryClient``), the bridge is a real mesh peer (hello / ping / provider
tables) instead of a hand-rolled socket loop, and ``GET /`` serves a dependency-free chat page instead of the SPA.
"""
from __future__ import annotations

import asyncio
import json
import math
import os
import threading
import time
from contextlib import asynccontextmanager
from typing import Any, AsyncIterator, Dict, List, Optional

from fastapi import FastAPI, Request
from fastapi.middleware.cors import CORSMiddleware
from fastapi.responses import HTMLResponse, JSONResponse, StreamingResponse

import logging

from .p2p import parse_join_link
from .p2p_runtime import P2PNode
from .utils.paths import bee2bee_home

log = logging.getLogger("bee2bee.gateway")
MODE = "fusion-inproc"
GEN_TIMEOUT_S = 90.0          # the reference bridge gives a generation 90 s (bridge.js:259-349)


class MetricsStore:
    """token / chat / visit counters persisted as one small JSON file (stands in for the `messages` table)."""

    def __init__(self, path: Optional[str] = None):
        self.path = path or str(bee2bee_home() / "gateway_metrics.json")
        self._lock = threading.Lock()
        self.data = {"tokens": 0, "chats": 0, "users": 0, "visits": 0}
        try:
            with open(self.path) as f:
                self.data.update({k: int(v) for k, v in json.load(f).items() if k in self.data})
        except (OSError, ValueError):
            pass

    def add(self, tokens: int = 0, chats: int = 0, visits: int = 0) -> None:
        with self._lock:
            self.data["tokens"] += max(0, int(tokens))
            self.data["chats"] += max(0, int(chats))
            self.data["visits"] += max(0, int(visits))
            try:
                os.makedirs(os.path.dirname(self.path), exist_ok=True)
                tmp = self.path + ".tmp"
                with open(tmp, "w") as f:
                    json.dump(self.data, f)
                os.replace(tmp, self.path)
            except OSError as e:      # metrics must never break a request
                log.warning("metrics persistence failed: %s", e)

    def snapshot(self) -> Dict[str, int]:
        with self._lock:
            return dict(self.data)


class MeshBridge:
    """The gateway's view of the mesh: one client-mode P2PNode, a pool of known node API addresses."""

    def __init__(self, seeds: Optional[List[str]] = None, node: Optional[P2PNode] = None, transport: Optional[str] = None):
        self.node = node
        self.transport = transport
        self._own_node = node is None
        self.seeds = [s for s in (seeds or []) if s]
        self.api_pool: Dict[str, Dict[str, Any]] = {}       # "host:port" of node HTTP sidecars -> last status

    async def start(self) -> None:
        if self.node is None:
            kw = {"transport": self.transport} if self.transport else {}
            self.node = P2PNode(host="127.0.0.1", port=0, region="Gateway", **kw)
            await self.node.start()
        for s in self.seeds:
            await self.connect_to_peer(s)

    async def stop(self) -> None:
        if self.node is not None and self._own_node:
            await self.node.stop()

    # ---------------------------------------------------------------- discovery
    async def connect_to_peer(self, addr: str) -> bool:
        try:
            if addr.startswith(("coithub", "p2pnet")):
                return bool(await self.node.connect_bootstrap(addr))
            return (await self.node._connect_peer(addr)) is not None
        except Exception as e:      # unreachable seeds are normal
            log.info("peer %s unreachable: %s", addr, e)
            return False

    async def register_join_link(self, link: str) -> Dict[str, Any]:
        info = parse_join_link(link)
        dialed = [a for a in info["bootstrap"] if await self.connect_to_peer(a)]
        return {"status": "registered" if dialed else "unreachable", "network": info["network"], "model": info["model"],
                "hash": info["hash"], "bootstrap": info["bootstrap"], "dialed": dialed}

    def stats(self) -> Dict[str, Any]:
        peers = self.node.peers if self.node else {}
        active = next(iter(peers.values()), None)
        return {"connected": bool(peers), "activeNode": (active or {}).get("addr"), "poolSize": len(peers) + len(self.api_pool),
                "peer_id": self.node.peer_id if self.node else None}

    def regional_mesh(self) -> Dict[str, List[Dict[str, Any]]]:
        """nodes grouped by region, the shape the dashboard renders (bridge.js getRegionalMesh)"""
        mesh: Dict[str, List[Dict[str, Any]]] = {}
        providers = {p["peer_id"]: p for p in self.node.list_providers()} if self.node else {}
        for pid, info in (self.node.peers if self.node else {}).items():
            prov = providers.get(pid, {})
            mesh.setdefault(info.get("region") or "Unknown", []).append({
                "peer_id": pid, "addr": info.get("addr"), "models": prov.get("models", []),
                "latency": info.get("last_pong_ms"), "status": info.get("health_status", "active"),
                "metrics": info.get("metrics", {}), "tag": prov.get("tag"),
                "api": (f"{info.get('api_host')}:{info.get('api_port')}" if info.get("api_port") else None)})
        return mesh

    # --------------------------------------------------------------- generation
    def _api_candidates(self, target: Optional[str]) -> List[str]:
        if target:
            return [target]
        out = []
        for info in (self.node.peers if self.node else {}).values():
            if info.get("api_port"):
                out.append(f"{info.get('api_host') or '127.0.0.1'}:{info['api_port']}")
        return out + [a for a in self.api_pool if a not in out]

    async def request(self, prompt: str, model: Optional[str], max_tokens: Optional[int], temperature: Optional[float],
                      target: Optional[str] = None) -> AsyncIterator[str]:
        """Yields text deltas.  Order of attempts mirrors the reference bridge: the node's HTTP sidecar
        (`POST /generate`, stream) first, the mesh `gen_request` second."""
        import httpx

        body = {"prompt": prompt, "stream": True, "temperature": 0.7 if temperature is None else temperature}
        if model and model != "default":
            body["model"] = model
        if max_tokens:
            body["max_new_tokens"] = int(max_tokens)
        for api in self._api_candidates(target):
            url = (api if api.startswith("http") else f"http://{api}") + "/generate"
            try:
                async with httpx.AsyncClient(timeout=httpx.Timeout(GEN_TIMEOUT_S, connect=2.0)) as cli:
                    async with cli.stream("POST", url, json=body) as resp:
                        if resp.status_code != 200:
                            continue
                        got = False
                        async for line in resp.aiter_lines():
                            piece = _ndjson_text(line)
                            if piece:
                                got = True
                                yield piece
                        if got:
                            return
            except Exception as e:
                log.info("direct generate via %s failed: %s", api, e)
        # mesh path: any provider of the model (cheapest, then fastest), else any provider at all
        node = self.node
        pid = None
        if model and model != "default":
            picked = node.pick_provider(model)
            pid = picked[0] if picked else None
        if pid is None:
            provs = [p for p in node.list_providers() if p["peer_id"] != node.peer_id]
            if not provs:
                raise RuntimeError("no_node_available")
            pid = provs[0]["peer_id"]
            model = model if model and model != "default" else (provs[0]["models"] or [None])[0]
        q: asyncio.Queue = asyncio.Queue()
        task = asyncio.create_task(node.request_generation(pid, prompt, max_new_tokens=int(max_tokens or 256),
                                                           model_name=model, temperature=body["temperature"], stream=True,
                                                           on_chunk=q.put_nowait, timeout=GEN_TIMEOUT_S))
        streamed = False
        while True:
            getter = asyncio.create_task(q.get())
            done, _ = await asyncio.wait({getter, task}, return_when=asyncio.FIRST_COMPLETED)
            if getter in done:
                streamed = True
                yield getter.result()
                continue
            getter.cancel()
            while not q.empty():
                streamed = True
                yield q.get_nowait()
            res = task.result()             # raises the provider's error, if any
            if not streamed and res.get("text"):
                yield res["text"]
            return


def _ndjson_text(line: str) -> str:
    """HF services stream NDJSON `{"text": ...}` lines, Ollama services raw text (services.py)."""
    if not line:
        return ""
    try:
        obj = json.loads(line)
    except ValueError:
        return line
    if isinstance(obj, dict):
        if obj.get("status") == "error":
            raise RuntimeError(obj.get("message", "provider error"))
        return "" if obj.get("done") else str(obj.get("text", ""))
    return line


_PAGE = """<!doctype html><meta charset=utf-8><title>bee2bee_b200 gateway</title>
<style>body{font:15px system-ui;margin:2rem auto;max-width:46rem}textarea{width:100%;height:5rem}
pre{white-space:pre-wrap;background:#f4f4f4;padding:1rem;min-height:4rem}</style>
<h2>bee2bee_b200 mesh gateway</h2><div id=st>...</div>
<p><input id=model placeholder="model (blank = any)"> <input id=tok type=number value=128 style="width:5rem"> tokens</p>
<textarea id=p placeholder="user: hello"></textarea><button onclick=go()>generate</button><pre id=out></pre>
<script>
async function st(){const r=await (await fetch('/api/p2p/status')).json();
 document.getElementById('st').textContent=`status: ${r.status} | peers: ${r.poolSize} | active: ${r.activeNode||'-'}`}
async function go(){const out=document.getElementById('out');out.textContent='';
 const r=await fetch('/api/p2p/generate',{method:'POST',headers:{'Content-Type':'application/json'},
  body:JSON.stringify({prompt:document.getElementById('p').value,model:document.getElementById('model').value||'default',
  max_tokens:+document.getElementById('tok').value})});
 const rd=r.body.getReader(),dec=new TextDecoder();for(;;){const {done,value}=await rd.read();if(done)break;
  out.textContent+=dec.decode(value)}}
st();setInterval(st,15000)</script>"""


def create_app(bridge: Optional[MeshBridge] = None, metrics: Optional[MetricsStore] = None) -> FastAPI:
    state: Dict[str, Any] = {"bridge": bridge, "metrics": metrics or MetricsStore()}

    @asynccontextmanager
    async def lifespan(app: FastAPI):
        if state["bridge"] is None:
            seeds = [s.strip() for s in os.environ.get("BEE2BEE_SEEDS", "").split(",") if s.strip()]
            state["bridge"] = MeshBridge(seeds)
        await state["bridge"].start()
        try:
            yield
        finally:
            await state["bridge"].stop()

    app = FastAPI(title="bee2bee_b200 gateway", lifespan=lifespan)
    app.add_middleware(CORSMiddleware, allow_origins=os.environ.get("CORS_ORIGINS", "*").split(","), allow_methods=["*"],
                       allow_headers=["*"])
    app.state.gateway = state

    @app.get("/", response_class=HTMLResponse)
    def page() -> str:
        """The web console (landing / quick-register / dashboard / chat, bee2bee_b200/web/index.html; the reference's
        React SPA, /root/reference/app/src/App.jsx, talks to the same /api/p2p/* routes).  ``?link=`` pre-fills the
        register form.  Falls back to the built-in minimal chat page when the static file is not installed."""
        state["metrics"].add(visits=1)
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "web", "index.html"), encoding="utf-8") as f:
                return f.read()
        except OSError:
            return _PAGE

    @app.post("/api/p2p/register")
    async def register(req: Request):
        body = await _json(req)
        link = body.get("link")
        if not link:
            return JSONResponse({"error": "Missing join link"}, status_code=400)
        try:
            res = await state["bridge"].register_join_link(link)
        except ValueError as e:
            return JSONResponse({"error": str(e)}, status_code=500)
        st = state["bridge"].stats()
        return {**res, "connected": st["connected"], "activeNode": st["activeNode"], "mode": MODE}

    @app.post("/api/p2p/generate")
    async def generate(req: Request):
        body = await _json(req)
        task = body.get("task") or {}
        prompt = task.get("prompt") or body.get("prompt")
        model = task.get("model") or body.get("model") or "default"
        target = task.get("targetNode") or body.get("targetNode")
        if not prompt:
            return JSONResponse({"error": "Prompt is required"}, status_code=400)
        br: MeshBridge = state["bridge"]

        async def relay():
            n_chars = 0
            yield " "                         # first byte right away, like the reference proxy
            try:
                async for piece in br.request(prompt, model, body.get("max_tokens"), body.get("temperature"), target):
                    n_chars += len(piece)
                    yield piece
            except Exception as e:
                yield f"\n\n[Error]: {e}"
            finally:
                if n_chars:
                    state["metrics"].add(tokens=math.ceil(n_chars / 4), chats=1)     # same estimate: chars / 4

        return StreamingResponse(relay(), media_type="text/event-stream",
                                 headers={"Cache-Control": "no-cache", "Connection": "keep-alive"})

    async def _status(target: Optional[str]) -> Dict[str, Any]:
        br: MeshBridge = state["bridge"]
        target_status = None
        if target:
            import httpx
            host = target if target.startswith("http") else f"http://{target}"
            try:
                async with httpx.AsyncClient(timeout=2.0) as cli:
                    r = await cli.get(host + "/")
                    if r.status_code == 200:
                        target_status = r.json()
                        br.api_pool[target] = {"seen": time.time(), **{k: target_status.get(k) for k in ("peer_id", "region", "models")}}
            except Exception as e:
                log.info("target %s unreachable: %s", target, e)
        st, mesh = br.stats(), br.regional_mesh()
        if target_status is not None:
            region = target_status.get("region") or "Local-Probe"
            nodes = mesh.setdefault(region, [])
            if not any(n.get("addr") == target or n.get("peer_id") == target_status.get("peer_id") for n in nodes):
                nodes.append({**target_status, "addr": target, "status": "active", "latency": 5, "tag": "direct-ingress"})
        active = st["connected"] or st["poolSize"] > 0 or target_status is not None
        return {**st, "mesh": mesh, "mode": MODE, "status": "active" if active else "idle"}

    @app.get("/api/p2p/status")
    async def status_get(target: Optional[str] = None):
        return await _status(target)

    @app.post("/api/p2p/status")
    async def status_post(req: Request, target: Optional[str] = None):
        body = await _json(req)
        peer = body.get("peer") or {}
        if body.get("action") == "discover_peer" and peer.get("addr"):
            asyncio.create_task(state["bridge"].connect_to_peer(peer["addr"]))
            return {"status": "discovery_initiated"}
        return await _status(target)

    @app.get("/api/p2p/global_metrics")
    def metrics_get():
        m = state["metrics"].snapshot()
        if not (m["tokens"] or m["chats"]):
            return {"visits": m["visits"], "chats": 0, "tokens": 0}
        return {"tokens": m["tokens"], "chats": m["chats"], "users": m["users"]}

    @app.post("/api/p2p/global_metrics")
    async def metrics_post(req: Request):
        tokens = int((await _json(req)).get("tokens") or 0)
        if tokens <= 0:
            return {"success": False}
        state["metrics"].add(tokens=tokens)
        return {"success": True}

    @app.exception_handler(404)
    async def not_found(req: Request, _exc):
        return JSONResponse({"error": f"Route {req.url.path} not found in the bee2bee_b200 mesh gateway"}, status_code=404)

    return app


async def _json(req: Request) -> Dict[str, Any]:
    try:
        body = await req.json()
        return body if isinstance(body, dict) else {}
    except Exception:
        return {}


def main(host: str = "0.0.0.0", port: Optional[int] = None, seeds: Optional[List[str]] = None) -> None:
    import uvicorn

    port = int(port or os.environ.get("API_PORT", 3000))
    uvicorn.run(create_app(MeshBridge(seeds) if seeds else None), host=host, port=port, log_level="warning")


if __name__ == "__main__":      # pragma: no cover
    main()
