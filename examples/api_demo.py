"""HTTP sidecar walk-through in one process (parity: /root/reference/examples/api_demo.py, extended to the routes that
actually generate): starts the FastAPI app with an embedded tiny model, then calls `/`, `/peers`, `/providers`,
`/generate` (buffered and streamed) and `/metrics` through the test client — no ports, no network.

    python examples/api_demo.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BEE2BEE_OFFLINE", "1")
os.environ.setdefault("BEE2BEE_TRANSPORT", "inproc")

from fastapi.testclient import TestClient  # noqa: E402

from bee2bee_b200 import api as api_mod  # noqa: E402
from bee2bee_b200.services import HFService  # noqa: E402


def main():
    api_mod.node = None
    with TestClient(api_mod.app) as c:                      # lifespan creates and starts a P2PNode
        svc = HFService("tiny-llama", 0.0)
        svc.load_sync()
        c.portal.call(api_mod.node.add_service, svc)        # announce the service on the node's loop
        print("GET /          ->", json.dumps(c.get("/").json())[:200])
        print("GET /peers     ->", c.get("/peers").json())
        print("GET /providers ->", c.get("/providers").json())
        r = c.post("/generate", json={"prompt": "user: hello", "max_new_tokens": 12, "temperature": 0.7}).json()
        print("POST /generate ->", {k: r[k] for k in ("status", "text", "metadata") if k in r})
        with c.stream("POST", "/generate", json={"prompt": "user: stream please", "max_new_tokens": 12, "stream": True}) as resp:
            print("POST /generate (stream) ->", [json.loads(l) for l in resp.iter_lines() if l][:4], "...")
        m = c.get("/metrics").json()
        print("GET /metrics   ->", {k: {kk: v[kk] for kk in ("requests", "tokens_generated", "kv_utilization")} for k, v in m.items()})
    api_mod.node = None


if __name__ == "__main__":
    main()
